"""filter3.cuh (BSP_FILTER_V3=1): the persistent prep -> link -> stream -> rescore form of the bound-and-rescore filter
(bulk-copied record double buffer, mbarriers, no CTA barriers) must return what filter_score_kernel and the oracle return,
bit for bit -- on the reference's own query fixtures (all three algorithms, kmer_skips 0 / 5: gap clauses, X / N masks,
reads with more rows than the 255-clause lanes take) and on adversarial small corpora (ties, repeats, exception cells)."""
import os

import numpy as np
import pytest

import common
import test_reference_reads as trr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("full_topn", [1, 0])
@pytest.mark.parametrize("skips", [0, 5])
@pytest.mark.parametrize("alg", ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"])
def test_filter3_matches_oracle_on_reference_reads(oracle_lib, monkeypatch, alg, skips, full_topn):
    fx = trr.load()
    entries, reads = trr.build_entries(fx)
    texts = [t for _, t in entries]
    seqs = [r[3] for r in reads]
    ox = common.build_oracle(oracle_lib, texts, 10, False)
    ro = ox.classify([s.encode() for s in seqs], skips=skips, algorithm=["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"].index(alg),
                     scoring=0, msm=0.5, threads=os.cpu_count() or 2)
    g = common.build_gpu(texts, 10, False, dict(kmer_skips=skips, query_algorithm=alg, scoring_algorithm="tfidf", full_topn=full_topn))
    try:
        monkeypatch.setenv("BSP_FILTER", "1")
        monkeypatch.setenv("BSP_FILTER_V3", "1")
        rg = g.classify_batch(seqs)
        common.assert_classify_equal(rg, ro, seqs, "filter3 %s skips=%d full_topn=%d" % (alg, skips, full_topn), hits_only=not full_topn)
        routes = g.last_routes(len(seqs))
        assert (routes[ro["status"] == 0] == 3).sum() >= 100          # (kmer_skips=5: reads beyond the gap-clause list take the positional scorer)
        monkeypatch.setenv("BSP_FILTER_V3", "0")
        r1 = g.classify_batch(seqs)
        for key in ("status", "label", "n_hits", "n_top"):
            assert np.array_equal(rg[key], r1[key]), key
    finally:
        g.close()


@pytest.mark.parametrize("seed", range(8))
def test_filter3_adversarial_corpora(oracle_lib, monkeypatch, seed):
    ents = common.make_corpus(900 + seed, n_entries=(3, 40), length=(20, 400), junk=0.01)
    reads = common.make_reads(seed, ents, n=96, max_len=160, n_frac=0.05)
    k = [4, 5, 6, 8][seed % 4]
    for alg in (0, 1, 2):
        ox = common.build_oracle(oracle_lib, ents, k, False)
        ro = ox.classify([r.encode() for r in reads], skips=seed % 3, algorithm=alg, scoring=0, msm=0.5, threads=2)
        g = common.build_gpu(ents, k, False, dict(kmer_skips=seed % 3, scoring_algorithm="tfidf", dense_tables=1, positional=1,
                                                  query_algorithm=["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]))
        try:
            monkeypatch.setenv("BSP_FILTER", "1")
            monkeypatch.setenv("BSP_FILTER_V3", "1")
            rg = g.classify_batch(reads)
            common.assert_classify_equal(rg, ro, reads, "filter3 adversarial seed=%d alg=%d" % (seed, alg))
        finally:
            g.close()
