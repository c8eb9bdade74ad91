"""GPU parity proper: CUDA path (through the C ABI) vs the C oracle on seeded adversarial corpora.
Bit-exact: postings, positions, norms, top-10 doc ids and score bit patterns, labels."""
import random

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

CASES = []
_rng = random.Random(7)
for _i in range(36):
    CASES.append(dict(seed=_i, k=_rng.choice([1, 2, 3, 3, 4, 5, 8, 10, 13, 16, 17, 24, 31]),
                      min_strand=_rng.random() < 0.4, skips=_rng.choice([0, 0, 0, 1, 2, 5]),
                      alg=_rng.choice([0, 1, 2]), sim=_rng.choice([0, 1]),
                      msm=_rng.choice([0.0, 0.5, 0.5, 0.9, 1.0]), junk=_rng.choice([0, 0, 0.01, 0.05])))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "s%d-k%d" % (c["seed"], c["k"]))
def test_index_and_classify_bit_exact(case, oracle_lib):
    ents = common.make_corpus(case["seed"], junk=case["junk"])
    reads = common.make_reads(1000 + case["seed"], ents, n=24)
    k, ms = case["k"], case["min_strand"]
    ox = common.build_oracle(oracle_lib, ents, k, ms)
    res_o = ox.classify([r.encode() for r in reads], skips=case["skips"], algorithm=case["alg"], scoring=case["sim"],
                        msm=case["msm"], threads=2)
    alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][case["alg"]]
    qkw = dict(kmer_skips=case["skips"], query_algorithm=alg_name, scoring_algorithm="bm25" if case["sim"] else "tfidf",
               query_term_min_should_match=case["msm"])
    # positional (general) path only
    g = common.build_gpu(ents, k, ms, dict(qkw, dense_tables=2, positional=1))
    try:
        common.assert_index_equal(g, ox)
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "positional")
        assert set(g.last_routes(len(reads))[res_o["status"] == 0]) <= {2}
    finally:
        g.close()
    # dense tables on (small k): reads whose clauses all have slop 2 take the dense route
    if k <= 8:
        g = common.build_gpu(ents, k, ms, dict(qkw, dense_tables=1, positional=1))
        try:
            res_g = g.classify_batch(reads)
            common.assert_classify_equal(res_g, res_o, reads, "dense+positional")
        finally:
            g.close()


def test_multi_segment_matches_single(oracle_lib, monkeypatch):
    """Tiny segments (BSP_SEG_TOKENS) force many positional segments and doc tiles."""
    monkeypatch.setenv("BSP_SEG_TOKENS", "1500")
    rng = random.Random(5)
    ents = [common.rand_seq(rng, rng.randint(200, 900)) for _ in range(40)]
    ents += [ents[3][:400] + common.rand_seq(rng, 100), "A" * 300, "ACG" * 150]
    reads = common.make_reads(77, ents, n=200, max_len=150)
    for k, alg, sim in ((4, 2, 0), (6, 1, 1), (9, 0, 0), (14, 2, 0)):
        ox = common.build_oracle(oracle_lib, ents, k, False)
        res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=alg, scoring=sim, msm=0.5, threads=4)
        alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]
        for dense in (2, 1):
            if dense == 1 and k > 8:
                continue
            g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name,
                                                      scoring_algorithm="bm25" if sim else "tfidf", dense_tables=dense,
                                                      positional=1))
            try:
                if dense == 2:
                    common.assert_index_equal(g, ox, sample_terms=[ox.term_at(i) for i in range(0, ox.stats()["n_terms"], 37)])
                res_g = g.classify_batch(reads)
                common.assert_classify_equal(res_g, res_o, reads, "k%d dense%d" % (k, dense))
                if dense == 1:
                    routes = g.last_routes(len(reads))
                    assert (routes == 1).sum() > 0
            finally:
                g.close()
