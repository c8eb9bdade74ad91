"""BASELINE.json configs as parity cases at (or near) their stated sizes.

C1 ("Local classify on CPU": 10 random 2 Mbp genomes, 10k reads of ~100 bp at 4 % error, default k, shipped
configuration.json) runs at FULL size against the C oracle, every read compared bit for bit, through each scorer.
The error-rate / read-length sweep of configs[4] runs on a smaller reference.  At C2 size the oracle cannot index the
reference in reasonable time; there the checks are the size-independent properties in bench.py (device path ==
host path, filter == exact-all scorer on a sample, source recovery)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1(oracle_lib):
    from biospectra_b200 import synth
    genomes, reads, src = synth.make_small_workload(n_genomes=10, genome_len=2_000_000, n_reads=10_000, R=100, err=0.04,
                                                    seed=1, read_seed=11)
    ox = oracle_lib.OracleIndex(10, False)
    for g in genomes:
        ox.add_entry(g)
    ox.finish()
    return genomes, reads, src, ox


def _gpu(genomes, **qkw):
    from biospectra_b200 import Configuration, Indexer
    ix = Indexer(Configuration(index_path="", kmer_size=10))
    for i, g in enumerate(genomes):
        ix.add_entry("%d.fna" % i, "syn|%d| synthetic genome %d" % (i, i), "", g)
    return ix.close_to_classifier(Configuration(index_path="", kmer_size=10, **qkw))


@pytest.mark.parametrize("alg", ["PAIRED_PROXIMITY", "NAIVE_KMER", "CHAIN_PROXIMITY"])
def test_c1_full_size_every_scorer(c1, alg, monkeypatch):
    genomes, reads, src, ox = c1
    a = {"NAIVE_KMER": 0, "CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}[alg]
    res_o = ox.classify(reads, skips=0, algorithm=a, scoring=0, msm=0.5, threads=8)
    assert (res_o["status"] == 0).all()
    # index statistics at full size
    g = _gpu(genomes, kmer_skips=0, query_algorithm=alg, scoring_algorithm="tfidf", full_topn=1, dense_tables=1, positional=1)
    try:
        assert g.stats() == ox.stats()
        for d in range(20):
            assert g.doc_info(d) == ox.doc_info(d)
        rng = np.random.default_rng(5)
        for t in rng.integers(0, 4 ** 10, 40):
            od, otf, opos = ox.postings(int(t))
            gd, gtf, gpos = g.postings(int(t))
            assert np.array_equal(od, gd) and np.array_equal(otf, gtf)
            assert np.array_equal(np.concatenate(opos) if opos else np.zeros(0, np.int32), gpos)
        for filt, expect_route in (("1", 3), ("0", 1)):
            monkeypatch.setenv("BSP_FILTER", filt)
            res_g = g.classify_batch(reads)
            common.assert_classify_equal(res_g, res_o, reads, "C1 %s filter=%s" % (alg, filt))
            assert (g.last_routes(len(reads)) == expect_route).all()
        monkeypatch.setenv("BSP_DISABLE_DENSE", "1")
        res_g = g.classify_batch(reads[:2000])
        common.assert_classify_equal(res_g, {k: v[:2000] for k, v in res_o.items()}, reads, "C1 %s positional" % alg)
        assert (g.last_routes(2000) == 2).all()
    finally:
        g.close()
    if alg == "PAIRED_PROXIMITY":
        # the reference's own sanity check (tools/verify_simulator_result.py:10-39): a classified read's top hit names
        # its source genome (forward strand: reads are sampled from the forward strand only)
        ok = res_o["label"] == 2
        assert (res_o["top_doc"][ok, 0] == 2 * np.asarray(src)[ok]).mean() > 0.99


def test_c1_result_set_mode_and_bm25(c1, monkeypatch):
    genomes, reads, src, ox = c1
    monkeypatch.setenv("BSP_FILTER", "1")
    res_o = ox.classify(reads, skips=0, algorithm=2, scoring=0, msm=0.5, threads=8)
    g = _gpu(genomes, kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf", dense_tables=1, positional=2)
    try:
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "C1 hits-only", hits_only=True)
        assert g.last_counters()["filter_reads"] == len(reads)
    finally:
        g.close()
    res_o = ox.classify(reads[:3000], skips=0, algorithm=2, scoring=1, msm=0.5, threads=8)
    g = _gpu(genomes, kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="bm25", full_topn=1, dense_tables=1,
             positional=2)
    try:
        common.assert_classify_equal(g.classify_batch(reads[:3000]), res_o, reads, "C1 bm25")
    finally:
        g.close()


@pytest.mark.parametrize("alg,scoring", [("PAIRED_PROXIMITY", "tfidf"), ("CHAIN_PROXIMITY", "tfidf"), ("PAIRED_PROXIMITY", "bm25")])
def test_c1_full_size_code_default_kmer_skips(c1, alg, scoring, monkeypatch):
    """C1 at full size with the CODE default kmer_skips = 5 (Configuration.java:37-44; the shipped json says 0): every phrase
    clause is a gap clause, evaluated by the gap join (gapjoin.cuh) and scored by the filter kernel / the exact kernel; all
    10,000 reads bit for bit against the oracle, and 2,000 of them through the positional scorer"""
    genomes, reads, src, ox = c1
    a = {"CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}[alg]
    res_o = ox.classify(reads, skips=5, algorithm=a, scoring=1 if scoring == "bm25" else 0, msm=0.5, threads=8)
    g = _gpu(genomes, kmer_skips=5, query_algorithm=alg, scoring_algorithm=scoring, full_topn=1, dense_tables=1, positional=1)
    try:
        for filt in ("1", "0"):
            monkeypatch.setenv("BSP_FILTER", filt)
            res_g = g.classify_batch(reads)
            common.assert_classify_equal(res_g, res_o, reads, "C1 skips=5 %s %s filter=%s" % (alg, scoring, filt))
            gj = g.last_gap_join()
            nc = int(res_o["n_clauses"][res_o["status"] == 0].sum())       # (PAIRED: a trailing Term clause is not a gap clause)
            assert nc - len(reads) <= gj["clauses"] <= nc and gj["matches"] > 0
            assert g.last_counters()["positional_reads"] == 0
        monkeypatch.setenv("BSP_DISABLE_GAP", "1")
        res_g = g.classify_batch(reads[:2000])
        common.assert_classify_equal(res_g, {k: v[:2000] for k, v in res_o.items()}, reads, "C1 skips=5 positional")
        assert (g.last_routes(2000)[res_o["status"][:2000] == 0] == 2).all()
    finally:
        g.close()


def test_error_rate_and_read_length_sweep(oracle_lib, monkeypatch):
    """configs[4]: 0-8 % error x 100-250 bp reads exercising the vague / unknown thresholds (smaller reference: 12 genomes
    of 300 kbp with two related pairs so that VAGUE ties occur), default code path (kmer_skips from the shipped json = 0)
    plus the code-default kmer_skips = 5 (positional path)."""
    from biospectra_b200 import synth
    rng = np.random.Generator(np.random.PCG64(77))
    genomes = [synth.genome_numpy(9, g, 300_000) for g in range(12)]
    genomes[5] = genomes[2].copy()
    genomes[5][rng.integers(0, 300_000, 3000)] = synth.ASCII[rng.integers(0, 4, 3000)]       # ~0.75 % diverged strain
    genomes[9] = genomes[2].copy()                                                           # identical strain: exact ties
    ox = oracle_lib.OracleIndex(10, False)
    for gg in genomes:
        ox.add_entry(gg.tobytes())
    ox.finish()
    monkeypatch.setenv("BSP_FILTER", "1")
    g = _gpu([x.tobytes() for x in genomes], kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf",
             full_topn=1, dense_tables=1, positional=1)
    g5 = _gpu([x.tobytes() for x in genomes], kmer_skips=5, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf",
              full_topn=1, dense_tables=1, positional=1)
    labels = np.zeros(3, np.int64)
    try:
        for err in (0.0, 0.02, 0.04, 0.08, 0.3):
            for R in (100, 150, 250):
                reads, _ = synth.sample_reads_numpy(genomes, 400, R, err, int(err * 1000) + R)
                reads += [b"N" * 40, b"ACGT" * 30, reads[0][:11], reads[1][:10]]
                res_o = ox.classify(reads, skips=0, algorithm=2, scoring=0, msm=0.5, threads=8)
                common.assert_classify_equal(g.classify_batch(reads), res_o, reads, "sweep e=%g R=%d" % (err, R))
                labels += np.bincount(res_o["label"][res_o["status"] == 0], minlength=3)
                res_o5 = ox.classify(reads[:120], skips=5, algorithm=2, scoring=0, msm=0.5, threads=8)
                common.assert_classify_equal(g5.classify_batch(reads[:120]), res_o5, reads, "sweep skips=5 e=%g R=%d" % (err, R))
    finally:
        g.close()
        g5.close()
    assert labels[0] > 0 and labels[1] > 0 and labels[2] > 0, labels       # UNKNOWN, VAGUE and CLASSIFIED all occur
