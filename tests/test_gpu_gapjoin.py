"""gapjoin.cuh: phrase clauses whose two k-mers are not adjacent in the read (every phrase clause with kmer_skips > 0 --
Configuration.java:37-44 defaults to 5, Classifier.java:160-200 builds slop = offset difference + 1 -- and the clauses next
to an N) evaluated for the whole batch against the positional index, sorted by first k-mer.  The CUDA path is compared
with the oracle bit for bit; the cases force the corners of the join: many small segments, first-term slices longer than
one staged block, second-term slices longer than one round, same-term clauses (periodic reads), the clause list and the
match tail growing, matches beyond the per-clause list cap."""
import random

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

ALGS = {"NAIVE_KMER": 0, "CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}


def _run(oracle_lib, ents, reads, k, skips, alg="PAIRED_PROXIMITY", scoring="tfidf", msm=0.5, min_strand=False, expect_join=True,
         what=""):
    ox = common.build_oracle(oracle_lib, ents, k, min_strand)
    ro = ox.classify([r.encode() for r in reads], skips=skips, algorithm=ALGS[alg], scoring=1 if scoring == "bm25" else 0, msm=msm,
                     threads=4)
    g = common.build_gpu(ents, k, min_strand, dict(kmer_skips=skips, query_algorithm=alg, scoring_algorithm=scoring,
                                                   query_term_min_should_match=msm, dense_tables=1, positional=1))
    try:
        rg = g.classify_batch(reads)
        common.assert_classify_equal(rg, ro, reads, what)
        gj = g.last_gap_join()
        if expect_join:
            assert gj["clauses"] > 0 and gj["chunks"] > 0, gj
        return g.last_routes(len(reads)), gj, ro
    finally:
        g.close()


def _genomes(seed, n, length, repeats=False):
    rng = random.Random(seed)
    ents = []
    for i in range(n):
        s = common.rand_seq(rng, length)
        if repeats and i % 2 == 0:                       # long homopolymer / short-period stretches: terms with many positions
            j = rng.randint(0, length // 2)
            s = s[:j] + "A" * 700 + s[j:j + 200] + "ACACACACACAC" * 60 + s[j + 200:]
        ents.append(s)
    return ents


def _sample_reads(seed, ents, n, length=100, err=0.03, n_frac=0.0):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        s = rng.choice(ents)
        a = rng.randint(0, max(0, len(s) - length))
        q = list(s[a:a + length])
        for i in range(len(q)):
            if rng.random() < err:
                q[i] = rng.choice("ACGT")
        if n_frac and rng.random() < n_frac:
            q[rng.randint(0, len(q) - 1)] = "N"
        out.append("".join(q))
    return out


@pytest.mark.parametrize("alg", ["PAIRED_PROXIMITY", "CHAIN_PROXIMITY"])
@pytest.mark.parametrize("skips", [1, 5])
def test_join_matches_oracle(oracle_lib, monkeypatch, alg, skips):
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(3, 12, 6000)
    reads = _sample_reads(4, ents, 300)
    routes, gj, ro = _run(oracle_lib, ents, reads, 6, skips, alg, what="join %s skips=%d" % (alg, skips))
    assert set(routes[ro["status"] == 0]) <= {1, 3}, "reads left the table scorers"          # no positional fallback needed
    assert gj["matches"] > 0


@pytest.mark.parametrize("scoring", ["tfidf", "bm25"])
def test_join_many_segments_and_small_k(oracle_lib, monkeypatch, scoring):
    monkeypatch.setenv("BSP_SEG_TOKENS", "3000")        # a segment per one or two documents
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(7, 20, 2500)
    reads = _sample_reads(8, ents, 200, length=80)
    _run(oracle_lib, ents, reads, 5, 5, scoring=scoring, what="many segments " + scoring)


def test_join_long_slices_and_same_term_clauses(oracle_lib, monkeypatch):
    """k = 4: 256 terms, every one with thousands of positions per segment (first-term slices span several staged blocks of
    512, second-term slices several rounds of 384); homopolymer and period-2 reads give clauses whose two k-mers are equal"""
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(11, 6, 30000, repeats=True)
    reads = _sample_reads(12, ents, 120, length=90)
    reads += ["A" * 60, "AC" * 40, "ACG" * 30, "A" * 30 + "C" * 30, ents[0][100:160] + "A" * 40]
    routes, gj, ro = _run(oracle_lib, ents, reads, 4, 5, what="long slices")
    routes, gj, ro = _run(oracle_lib, ents, reads, 4, 2, alg="CHAIN_PROXIMITY", what="long slices chain")


def test_join_reads_with_n_keep_the_table_rows(oracle_lib, monkeypatch):
    """kmer_skips = 0: only the clauses next to an N are gap clauses; the others keep their table rows"""
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(21, 10, 5000)
    reads = _sample_reads(22, ents, 300, n_frac=0.5)
    routes, gj, ro = _run(oracle_lib, ents, reads, 6, 0, what="reads with N")
    assert 0 < gj["clauses"] < 3 * len(reads)


def test_clause_list_grows(oracle_lib, monkeypatch):
    """BSP_GAP_CAP=2 and kmer_skips = 0 -> 5 on the open index: the estimate of the clause count is what sizes the list;
    with reads far longer than the estimate assumes nothing breaks either (phase 1 runs again with a longer list)"""
    monkeypatch.setenv("BSP_GAP_CAP", "2")
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(31, 8, 4000)
    reads = _sample_reads(32, ents, 300)
    ox = common.build_oracle(oracle_lib, ents, 6, False)
    ro = ox.classify([r.encode() for r in reads], skips=5, algorithm=1, scoring=0, msm=0.5, threads=4)
    g = common.build_gpu(ents, 6, False, dict(kmer_skips=0, query_algorithm="CHAIN_PROXIMITY", dense_tables=1, positional=1))
    try:
        from biospectra_b200 import Configuration
        g.set_query(Configuration(index_path="", kmer_size=6, kmer_skips=5, query_algorithm="CHAIN_PROXIMITY", full_topn=1))
        rg = g.classify_batch(reads)
        common.assert_classify_equal(rg, ro, reads, "clause list")
        assert g.last_gap_join()["clauses"] == int(ro["n_clauses"][ro["status"] == 0].sum())       # (CHAIN: every clause is a phrase)
    finally:
        g.close()


def test_match_tail_overflow_then_grows(oracle_lib, monkeypatch):
    """kmer_skips = 0, half the reads carry an N, BSP_GAP_CAP=2: two listed clauses at first (phase 1 runs again), and the
    match tail behind the exception arrays (2 x 64 entries) is too short for the batch's matches -- every read with a gap
    clause falls back to the positional scorer; the next batch finds the room it asked for"""
    monkeypatch.setenv("BSP_GAP_CAP", "2")
    monkeypatch.setenv("BSP_FILTER", "1")
    ents = _genomes(33, 8, 4000)
    ents += [ents[0][:3000] + common.rand_seq(random.Random(1), 500)] * 5         # near-identical documents: several matches per clause
    reads = _sample_reads(34, ents, 400, n_frac=0.5)
    ox = common.build_oracle(oracle_lib, ents, 6, False)
    ro = ox.classify([r.encode() for r in reads], skips=0, algorithm=2, scoring=0, msm=0.5, threads=4)
    g = common.build_gpu(ents, 6, False, dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", dense_tables=1, positional=1))
    try:
        rg = g.classify_batch(reads)
        common.assert_classify_equal(rg, ro, reads, "first batch (tail too short)")
        first = g.last_counters()
        assert first["positional_reads"] > 0 and g.last_gap_join()["matches"] == 0
        rg = g.classify_batch(reads)
        common.assert_classify_equal(rg, ro, reads, "second batch")
        second = g.last_counters()
        assert second["positional_reads"] < first["positional_reads"] and g.last_gap_join()["matches"] > 128
    finally:
        g.close()


def test_more_matches_than_the_list_cap(oracle_lib, monkeypatch):
    """80 copies of one document: every clause of a read drawn from it matches 80+ documents, more than the 64 a gap
    clause's list keeps -- those reads must take the positional scorer and still agree with the oracle"""
    monkeypatch.setenv("BSP_FILTER", "1")
    rng = random.Random(41)
    base = common.rand_seq(rng, 1500)
    ents = [base] * 80 + _genomes(42, 4, 1500)
    reads = _sample_reads(43, [base], 40, err=0.0) + _sample_reads(44, ents[80:], 40)
    routes, gj, ro = _run(oracle_lib, ents, reads, 7, 5, what="list cap")
    assert (routes[:40] == 2).all()
