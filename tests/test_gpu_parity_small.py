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
def test_index_and_classify_bit_exact(case, oracle_lib, monkeypatch):
    ents = common.make_corpus(case["seed"], junk=case["junk"])
    reads = common.make_reads(1000 + case["seed"], ents, n=24)
    k, ms = case["k"], case["min_strand"]
    ox = common.build_oracle(oracle_lib, ents, k, ms)
    res_o = ox.classify([r.encode() for r in reads], skips=case["skips"], algorithm=case["alg"], scoring=case["sim"],
                        msm=case["msm"], threads=2)
    alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][case["alg"]]
    qkw = dict(kmer_skips=case["skips"], query_algorithm=alg_name, scoring_algorithm="bm25" if case["sim"] else "tfidf",
               query_term_min_should_match=case["msm"])
    # positional (general) path only: warp-per-(segment, read) kernel (few docs per segment) and the block kernel
    g = common.build_gpu(ents, k, ms, dict(qkw, dense_tables=2, positional=1))
    try:
        common.assert_index_equal(g, ox)
        for block_kernel in ("0", "1"):
            monkeypatch.setenv("BSP_POS_BLOCK_KERNEL", block_kernel)
            res_g = g.classify_batch(reads)
            common.assert_classify_equal(res_g, res_o, reads, "positional block_kernel=" + block_kernel)
            assert set(g.last_routes(len(reads))[res_o["status"] == 0]) <= {2}
        monkeypatch.delenv("BSP_POS_BLOCK_KERNEL")
    finally:
        g.close()
    # dense 4-bit tables on (small k): reads whose clauses all have slop 2 take a dense route --
    # every doc scored exactly (BSP_FILTER=0), or bound-and-rescore (BSP_FILTER=1; bm25 through per-norm-class LUTs)
    if k <= 8:
        for filt in ("0", "1"):
            monkeypatch.setenv("BSP_FILTER", filt)
            g = common.build_gpu(ents, k, ms, dict(qkw, dense_tables=1, positional=1))
            try:
                res_g = g.classify_batch(reads)
                common.assert_classify_equal(res_g, res_o, reads, "dense+positional filter=" + filt)
                routes = set(g.last_routes(len(reads))[res_o["status"] == 0])
                assert routes <= ({1, 2} if filt == "0" else {1, 2, 3}), routes
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
        for dense, filt in ((2, "0"), (1, "0"), (1, "1")):
            if dense == 1 and k > 8:
                continue
            monkeypatch.setenv("BSP_FILTER", filt)
            g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name,
                                                      scoring_algorithm="bm25" if sim else "tfidf", dense_tables=dense,
                                                      positional=1))
            try:
                if dense == 2:
                    common.assert_index_equal(g, ox, sample_terms=[ox.term_at(i) for i in range(0, ox.stats()["n_terms"], 37)])
                res_g = g.classify_batch(reads)
                common.assert_classify_equal(res_g, res_o, reads, "k%d dense%d filter%s" % (k, dense, filt))
                if dense == 2:
                    monkeypatch.setenv("BSP_POS_BLOCK_KERNEL", "1")
                    common.assert_classify_equal(g.classify_batch(reads), res_o, reads, "k%d block kernel" % k)
                    monkeypatch.delenv("BSP_POS_BLOCK_KERNEL")
                if dense == 1:
                    routes = g.last_routes(len(reads))
                    assert ((routes == 1) | (routes == 3)).sum() > 0
                    if filt == "1" and not sim:
                        assert (routes == 3).sum() > 0
            finally:
                g.close()


def _many_docs_corpus(seed, n_docs, length, k):
    """many short related docs: a few ancestor sequences, each doc a mutated copy (so scores are close and ties,
    repeats and exception cells occur), plus homopolymer / periodic docs"""
    rng = random.Random(seed)
    anc = [common.rand_seq(rng, length) for _ in range(6)]
    ents = []
    for i in range(n_docs):
        mode = rng.random()
        if mode < 0.04:
            ents.append(rng.choice("ACGT") * rng.randint(k + 2, length))
        elif mode < 0.08:
            ents.append((common.rand_seq(rng, rng.randint(2, 5)) * length)[:length])
        else:
            a = list(rng.choice(anc))
            for j in range(len(a)):
                if rng.random() < 0.03:
                    a[j] = rng.choice("ACGT")
            lo = rng.randint(0, length // 3)
            ents.append("".join(a[lo:lo + rng.randint(length // 2, length)]))
    return ents


@pytest.mark.parametrize("k,alg,n_docs", [(5, 2, 700), (4, 1, 1300), (6, 0, 500), (5, 2, 90), (5, 2, 2600)])
def test_filter_path_many_docs(oracle_lib, monkeypatch, k, alg, n_docs):
    """The bound-and-rescore kernel with CTAs of several warps (>= 10 threads hold valid docs, so the 10th-best
    lower bound really prunes), forced candidates from exception cells, and candidate overflow (many exact ties)
    falling back to the exact kernel."""
    ents = _many_docs_corpus(40 + k, n_docs // 2, 260, k)          # 2 docs per entry
    reads = common.make_reads(900 + k, ents, n=160, max_len=140, n_frac=0.0)
    ox = common.build_oracle(oracle_lib, ents, k, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=alg, scoring=0, msm=0.5, threads=4)
    alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]
    monkeypatch.setenv("BSP_FILTER", "1")
    g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name, scoring_algorithm="tfidf",
                                              dense_tables=1, positional=2))
    try:
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "filter k%d alg%d" % (k, alg))
        cnt = g.last_counters()
        assert cnt["filter_reads"] > 0, cnt
        assert cnt["filter_candidates"] > 0
    finally:
        g.close()
    # several doc tiles per read (CTAs of 32 threads = 1,024 docs each): per-tile lists merged by merge_kernel
    monkeypatch.setenv("BSP_FILTER_THREADS", "32")
    for full in (1, 0):
        g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name, scoring_algorithm="tfidf",
                                                  dense_tables=1, positional=2, full_topn=full))
        try:
            common.assert_classify_equal(g.classify_batch(reads), res_o, reads, "filter tiles k%d alg%d" % (k, alg), hits_only=not full)
            assert g.last_counters()["filter_reads"] > 0
        finally:
            g.close()
    monkeypatch.delenv("BSP_FILTER_THREADS")
    # result-set mode (the default): only the hits equal to the maximum are reported, from fewer rescored docs
    g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name, scoring_algorithm="tfidf",
                                              dense_tables=1, positional=2, full_topn=0))
    try:
        res_h = g.classify_batch(reads)
        common.assert_classify_equal(res_h, res_o, reads, "filter hits-only k%d alg%d" % (k, alg), hits_only=True)
        cnt_h = g.last_counters()
        assert cnt_h["filter_reads"] > 0 and cnt_h["filter_candidates"] <= cnt["filter_candidates"]
    finally:
        g.close()


def test_filter_overflow_falls_back(oracle_lib, monkeypatch):
    """> 128 docs tie exactly at the top: the filter cannot keep the candidate list and hands the read to the
    exact kernel; results stay identical."""
    rng = random.Random(9)
    core = common.rand_seq(rng, 120)
    ents = [core for _ in range(150)] + [common.rand_seq(rng, 120) for _ in range(40)]
    reads = [core[5:95], core[20:80], common.rand_seq(rng, 80)] + [e[3:90] for e in ents[150:160]]
    ox = common.build_oracle(oracle_lib, ents, 5, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=2, scoring=0, msm=0.5, threads=2)
    monkeypatch.setenv("BSP_FILTER", "1")
    g = common.build_gpu(ents, 5, False, dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf",
                                              dense_tables=1, positional=2))
    try:
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "overflow")
        cnt = g.last_counters()
        assert cnt["filter_overflow_reads"] >= 2, cnt
    finally:
        g.close()


@pytest.mark.parametrize("alg", [0, 1, 2])
def test_filter_long_reads_up_to_255_clauses(oracle_lib, monkeypatch, alg):
    """Reads with up to 255 clauses stay on the packed-lane filter (u8 match counters, u16 weight lanes); longer ones are
    handed to the exact scorer.  Both must agree with the oracle."""
    ents = _many_docs_corpus(77, 300, 700, 5)
    rng = random.Random(5)
    reads = []
    for i in range(120):
        e = rng.choice(ents)
        ln = rng.choice([150, 200, 254, 258, 259, 262, 300, 420, 520, 600])
        a = rng.randint(0, max(0, len(e) - 10))
        s = (e[a:] + common.rand_seq(rng, 700))[:ln]
        reads.append("".join(c if rng.random() > 0.03 else rng.choice("ACGT") for c in s))
    ox = common.build_oracle(oracle_lib, ents, 5, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=alg, scoring=0, msm=0.5, threads=4)
    monkeypatch.setenv("BSP_FILTER", "1")
    g = common.build_gpu(ents, 5, False, dict(kmer_skips=0, query_algorithm=["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg],
                                              scoring_algorithm="tfidf", dense_tables=1, positional=2))
    try:
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "long reads alg%d" % alg)
        routes = g.last_routes(len(reads))
        ncl = res_o["n_clauses"]
        assert (ncl[routes == 3] <= 255).all() and (routes[ncl > 255] == 1).all()       # (a lane overflow also hands over)
        assert ((routes == 3) & (ncl > 181)).sum() > 0 and (routes == 1).sum() > 0
    finally:
        g.close()


def test_edge_batches_and_clause_limits(oracle_lib, monkeypatch):
    """Empty batch, reads shorter than k, exactly 2 tokens, and the reference's 10,000-clause limit
    (Classifier.java:110 BooleanQuery.setMaxClauseCount(10000): longer reads throw and are skipped); reads with
    thousands of clauses go through the multi-stage loops of the exact and positional scorers."""
    rng = random.Random(12)
    ents = [common.rand_seq(rng, rng.randint(3000, 12000)) for _ in range(9)] + ["ACGT" * 3000]
    long_src = ents[2] + ents[5]
    reads = ["ACGTA", "ACGTAC", "ACGTACG", ents[0][:5], ents[1][10:2000], ents[2][:9000], (long_src * 2)[:10004],
             (long_src * 2)[:10005], (long_src * 2)[:10006], (long_src * 2)[:14000], "ACGT" * 2600, "N" * 30]
    k = 6
    ox = common.build_oracle(oracle_lib, ents, k, False)
    for alg, sim in ((0, 0), (1, 0), (2, 1), (2, 0)):
        res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=alg, scoring=sim, msm=0.5, threads=4)
        assert res_o["status"][0] == 1 and res_o["status"][1] == 1 and res_o["status"][2] == 0
        alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]
        if alg == 0:
            # NAIVE: n_clauses = n_tokens = len - k + 1: 10,005 bases -> 10,000 clauses (kept), 10,006 -> 10,001 (dropped)
            assert res_o["n_clauses"][7] == 10000 and res_o["status"][7] == 0 and res_o["status"][8] == 1
        for dense in (1, 2):
            monkeypatch.setenv("BSP_FILTER", "1")
            g = common.build_gpu(ents, k, False, dict(kmer_skips=0, query_algorithm=alg_name, scoring_algorithm="bm25" if sim else "tfidf",
                                                      dense_tables=dense, positional=1))
            try:
                res_g = g.classify_batch(reads)
                common.assert_classify_equal(res_g, res_o, reads, "edge alg%d sim%d dense%d" % (alg, sim, dense))
                empty = g.classify_packed(np.zeros(0, np.uint8), np.zeros(1, np.int64))
                assert len(empty["status"]) == 0
                with pytest.raises(ValueError):
                    g.classify_batch(["ACGTACGT", ""])          # Classifier.java:342-344: null/empty sequence
            finally:
                g.close()


@pytest.mark.parametrize("alg,skips", [(2, 0), (1, 2), (2, 5)])
def test_positional_warp_kernel_staged_and_direct_paths(oracle_lib, monkeypatch, alg, skips):
    """The warp-per-(segment, read) scorer stages a phrase clause in shared memory when both terms' slices fit
    (<= 128 postings, <= 384 positions per term and segment) and walks global memory otherwise.  200 docs that share a
    stretch (terms with df = 200 > 128), one periodic doc (a term with hundreds of positions) and unrelated docs put
    clauses on both sides of both caps; every read is compared bit for bit."""
    rng = random.Random(91)
    shared = common.rand_seq(rng, 60)
    ents = []
    for i in range(100):
        ents.append(common.rand_seq(rng, rng.randint(20, 120)) + shared + common.rand_seq(rng, rng.randint(0, 80)))
    ents[7] = ("ACGGTC" * 400) + shared                      # ~400 positions per term of the period
    ents[8] = "A" * 500 + shared[:30]                        # same-term phrases (repeat path)
    k = 5
    reads = common.make_reads(1091, ents, n=96, max_len=140, n_frac=0.1)
    reads += [shared[5:55], ("ACGGTC" * 20), "A" * 40, shared[:20] + "N" + shared[21:50]]
    ox = common.build_oracle(oracle_lib, ents, k, False)
    res_o = ox.classify([r.encode() for r in reads], skips=skips, algorithm=alg, scoring=0, msm=0.3, threads=2)
    alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]
    g = common.build_gpu(ents, k, False, dict(kmer_skips=skips, query_algorithm=alg_name, query_term_min_should_match=0.3,
                                              dense_tables=2, positional=1))
    try:
        monkeypatch.setenv("BSP_POS_BLOCK_KERNEL", "0")
        res_g = g.classify_batch(reads)
        common.assert_classify_equal(res_g, res_o, reads, "warp positional alg=%d skips=%d" % (alg, skips))
        assert set(g.last_routes(len(reads))[res_o["status"] == 0]) <= {2}
    finally:
        g.close()


def test_pipeline_slots_with_tiny_sub_batches(oracle_lib, monkeypatch):
    """bsp_classify_packed runs sub-batches through three pipeline slots (phase 1 of sub-batch i, scorer launch of i-1,
    completion of i-2 per iteration).  Sub-batches of 1, 2, 3, 5 and 7 reads over 61 reads exercise every fill / drain
    shape and the reuse of every slot; each setting must reproduce the oracle bit for bit."""
    ents = common.make_corpus(77, n_entries=(6, 7), length=(150, 300))
    reads = common.make_reads(1077, ents, n=61, n_frac=0.3)
    ox = common.build_oracle(oracle_lib, ents, 5, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=2, scoring=0, msm=0.4, threads=2)
    g = common.build_gpu(ents, 5, False, dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", query_term_min_should_match=0.4))
    try:
        for sub in ("1", "2", "3", "5", "7", "64"):
            monkeypatch.setenv("BSP_SUB_BATCH", sub)
            res_g = g.classify_batch(reads)
            common.assert_classify_equal(res_g, res_o, reads, "sub-batch " + sub)
            routes = g.last_routes(len(reads))
            assert (routes[res_o["status"] == 0] > 0).all() and (routes[res_o["status"] != 0] == 0).all()
    finally:
        g.close()
