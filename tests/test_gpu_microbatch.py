"""bsp_classify_one: the blocking single-read entry point behind Classifier.classify(header, sequence)
(classify/Classifier.java:341-366), called from many threads at once like LocalClassifier's pool
(LocalClassifier.java:96-118).  Concurrent calls must coalesce into GPU micro-batches and every caller must get
exactly its own read's result, bit-identical to the oracle."""
import threading

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _stack(rows):
    out = {k: np.array([r[k] for r in rows], np.int32) for k in ("status", "label", "n_hits", "n_top")}
    out["top_doc"] = np.stack([r["top_doc"] for r in rows])
    out["top_score"] = np.stack([r["top_score"] for r in rows])
    return out


def _run_threads(g, reads, n_threads):
    rows = [None] * len(reads)
    errors = []

    def work(t):
        try:
            for i in range(t, len(reads), n_threads):
                rows[i] = g.classify_one(reads[i].encode())
        except Exception as e:          # pragma: no cover - reported below
            errors.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    return _stack(rows)


@pytest.mark.parametrize("alg", ["PAIRED_PROXIMITY", "NAIVE_KMER"])
def test_concurrent_single_reads_match_oracle_and_coalesce(alg, oracle_lib):
    ents = common.make_corpus(41, n_entries=(5, 6), length=(200, 400))
    reads = common.make_reads(1041, ents, n=384)
    ox = common.build_oracle(oracle_lib, ents, 6, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"].index(alg),
                        scoring=0, msm=0.5, threads=2)
    g = common.build_gpu(ents, 6, False, dict(kmer_skips=0, query_algorithm=alg, query_term_min_should_match=0.5))
    try:
        g.set_microbatch(4096, 2000)
        res = _run_threads(g, reads, 32)
        common.assert_classify_equal(res, res_o, reads, "micro-batched " + alg)
        st = g.microbatch_stats()
        assert st["reads"] == len(reads)
        assert st["batches"] < len(reads) // 4, st           # 32 callers at a time: far fewer GPU batches than reads
        # a lone caller is its own batch; the window only adds latency
        g.set_microbatch(4096, 0)
        one = g.classify_one(reads[0].encode())
        common.assert_classify_equal(_stack([one]), {k: v[:1] for k, v in res_o.items()}, reads[:1], "single")
        assert g.microbatch_stats()["batches"] == st["batches"] + 1
        # max_reads = 1 switches coalescing off: one GPU batch per read, same answers
        g.set_microbatch(1, 0)
        res1 = _run_threads(g, reads[:64], 8)
        common.assert_classify_equal(res1, {k: v[:64] for k, v in res_o.items()}, reads[:64], "max_reads=1")
    finally:
        g.close()


def test_single_read_argument_errors_and_dropped_reads(oracle_lib):
    ents = common.make_corpus(42, n_entries=(3, 4), length=(100, 200))
    g = common.build_gpu(ents, 8, False, dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY"))
    try:
        with pytest.raises(ValueError):
            g.classify_one(b"")                                   # Classifier.java:342-344
        with pytest.raises(Exception):
            g.set_microbatch(0, 10)
        r = g.classify_one(b"ACGTACGT")                            # one token: the reference's query is null -> read dropped
        assert r["status"] == 1 and r["n_hits"] == 0
        r = g.classify_one(b"NNNNNNNNNNNNNNNNNNNN")
        assert r["status"] == 1
    finally:
        g.close()
