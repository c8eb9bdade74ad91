"""The reference's own query fixtures (benchmark/bacteria_ncbi_simulated/sample_*.fa, benchmark/simHC.20.500/*.fna; subset
committed as tests/golden/reference_reads.json.gz by tests/golden/make_reference_reads.py) through the oracles and the CUDA
path.  The genomes the reads were drawn from are not in the reference repository, so the index is synthetic and built from
the reads themselves: one entry per source genome header (simulated reads) / per FAMeS file (simHC reads), each the
concatenation of its reads with random spacers -- every read then has a true source entry, X/N-masked Sanger reads meet
X/N-masked reference text, and the 1,791 bp reads produce > 510 clauses (past the filter kernel's 255-clause lanes).

What is checked: the two oracles agree (CPU); the CUDA path returns the oracle's hit lists bit for bit for all three query
algorithms, kmer_skips 0 and the code default 5, tf-idf and bm25, with and without the filter kernel (GPU); and the
statement of tools/verify_simulator_result.py:10-39 -- a classified read's top hit names the read's source genome -- holds
for the error-free sample_0.0.fa reads."""
import gzip
import json
import os
import random

import numpy as np
import pytest

import common

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with gzip.open(os.path.join(HERE, "golden", "reference_reads.json.gz")) as f:
        return json.loads(f.read().decode())


def build_entries(fx, seed=9):
    """-> (entries [(name, text)], reads [(kind, file, header, seq, entry index)])"""
    rng = random.Random(seed)
    groups, order = {}, []
    for kind in ("simulated", "simhc"):
        for r in fx[kind]:
            key = r["header"] if kind == "simulated" else r["file"]
            if key not in groups:
                groups[key] = []
                order.append(key)
            groups[key].append((kind, r))
    entries, reads = [], []
    for e, key in enumerate(order):
        parts = []
        for kind, r in groups[key]:
            parts.append(common.rand_seq(rng, rng.randint(20, 60)))
            parts.append(r["seq"])
            reads.append((kind, r["file"], r["header"], r["seq"], e))
        parts.append(common.rand_seq(rng, 40))
        entries.append((key, "".join(parts)))
    return entries, reads


def test_fixture_is_the_reference_subset():
    fx = load()
    assert len(fx["simulated"]) == 800 and len(fx["simhc"]) == 320
    assert len({r["file"] for r in fx["simulated"]}) == 20 and len({r["file"] for r in fx["simhc"]}) == 20
    lens = [len(r["seq"]) for r in fx["simulated"]]
    assert 90 <= min(lens) and max(lens) <= 109                       # RandomReadSampler: R +- 10 %
    hl = [len(r["seq"]) for r in fx["simhc"]]
    assert max(hl) == 1791 and min(hl) >= 155
    assert any("X" in r["seq"] for r in fx["simhc"]) and any("N" in r["seq"] for r in fx["simhc"])
    assert all(r["seq"] == r["seq"].upper().strip() for r in fx["simhc"])    # FASTAReader semantics already applied


def test_oracles_agree_on_reference_reads(oracle_lib):
    import spec_oracle as so
    fx = load()
    entries, reads = build_entries(fx)
    # a small slice for the literal Python oracle: entries of the first simulated sources and one simHC file
    sub_e = [0, 1, 2, len(entries) - 1]
    texts = [entries[e][1][:3000] for e in sub_e]
    pick = [r for r in reads if r[4] in sub_e][:6] + [r for r in reads if r[0] == "simhc" and r[4] == sub_e[-1]][:3]
    seqs = [r[3][:400] for r in pick]
    k = 10
    six = so.Index(k, False)
    for i, t in enumerate(texts):
        six.add_entry("f%d.fa" % i, "h%d" % i, "", t)
    ox = common.build_oracle(oracle_lib, texts, k, False)
    for alg in (0, 1, 2):
        for skips in (0, 5):
            rc = ox.classify([s.encode() for s in seqs], skips=skips, algorithm=alg, scoring=0, msm=0.5, threads=1)
            for i, s in enumerate(seqs):
                rs = so.search(six, s, skips=skips, algorithm=alg, scoring=so.TFIDF, msm_ratio=0.5)
                if rs is None:
                    assert rc["status"][i] == 1
                    continue
                hits = rs[0]
                nt = int(rc["n_top"][i])
                assert [d for d, _ in hits] == rc["top_doc"][i, :nt].tolist(), (alg, skips, i)
                assert [int(np.float32(x).view(np.uint32)) for _, x in hits] == rc["top_score"][i, :nt].view(np.uint32).tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("scoring", ["tfidf", "bm25"])
@pytest.mark.parametrize("skips", [0, 5])
@pytest.mark.parametrize("alg", ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"])
def test_gpu_matches_oracle_on_reference_reads(oracle_lib, alg, skips, scoring):
    fx = load()
    entries, reads = build_entries(fx)
    texts = [t for _, t in entries]
    seqs = [r[3] for r in reads]
    k = 10
    ox = common.build_oracle(oracle_lib, texts, k, False)
    ro = ox.classify([s.encode() for s in seqs], skips=skips, algorithm=["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"].index(alg),
                     scoring=1 if scoring == "bm25" else 0, msm=0.5, threads=os.cpu_count() or 2)
    g = common.build_gpu(texts, k, False, dict(kmer_skips=skips, query_algorithm=alg, scoring_algorithm=scoring))
    try:
        assert g.stats() == ox.stats()
        for filt in ("1", "0"):
            os.environ["BSP_FILTER"] = filt
            try:
                rg = g.classify_batch(seqs)
            finally:
                del os.environ["BSP_FILTER"]
            common.assert_classify_equal(rg, ro, seqs, "reference reads %s skips=%d %s filter=%s" % (alg, skips, scoring, filt))
        routes = g.last_routes(len(seqs))
        long_reads = np.array([len(s) > 1100 for s in seqs])
        assert long_reads.any() and (ro["status"][long_reads] == 0).all()
        if skips == 0 and scoring == "tfidf":
            # tools/verify_simulator_result.py:10-39: the top hit of an error-free simulated read is its source genome
            for i, r in enumerate(reads):
                if r[0] == "simulated" and r[1] == "sample_0.0.fa" and ro["status"][i] == 0:
                    tops = rg["top_doc"][i, :int(rg["n_hits"][i])].tolist()
                    assert 2 * r[4] in tops, (i, r[2], tops)                       # forward doc of the source entry
                    assert g.doc_fields(2 * r[4])[1] == "h%d" % r[4]
        assert set(np.unique(routes)) <= {0, 1, 2, 3}
    finally:
        g.close()
