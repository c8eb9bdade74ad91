"""BASELINE configs[1] at FULL size (3,000 synthetic genomes, 10 Gbp, 6,000 documents, k = 10, shipped
configuration.json).  The CPU oracle cannot index 10 Gbp in test time, so parity is checked through size-independent
properties: three independent scorers of the library (packed-lane filter + exact rescoring, every-doc-exact fp64
scorer over the 4-bit tables, positional postings with the literal sloppy-phrase sweep) must agree bit for bit, batches
must be splittable, and the device-pointer and host-buffer entry points must agree.  (Each scorer is checked against
the oracle at small and C1 size in the other test files; bench.py checks 200,000 reads against the oracle on a sample
index.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import torch
    from biospectra_b200 import Configuration, Indexer, synth
    free, _total = torch.cuda.mem_get_info(0)
    if free < 170 * (1 << 30):
        pytest.skip("needs a whole B200 (170 GB free)")
    dev = torch.device("cuda", 0)
    lengths = synth.genome_lengths(3000, 10_000_000_000, 2)
    shipped = dict(index_path="", kmer_size=10, kmer_skips=0, min_strand_kmer=False, query_term_min_should_match=0.5,
                   query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf")
    ix = Indexer(Configuration(**shipped))
    data = synth.build_reference_and_reads_cuda(ix, lengths, 40_000, 100, 0.04, 2, 12, dev, n_frac=0.01)
    torch.cuda.empty_cache()
    clf = ix.close_to_classifier(Configuration(full_topn=1, **shipped))
    yield clf, data
    clf.close()


def _same(a, b, n_top_key="n_top"):
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["n_hits"], b["n_hits"]) and np.array_equal(a["label"], b["label"])
    assert np.array_equal(a[n_top_key], b[n_top_key])
    live = np.arange(10)[None, :] < a[n_top_key][:, None]
    assert np.array_equal(a["top_doc"][live], b["top_doc"][live])
    assert np.array_equal(a["top_score"][live].view(np.uint32), b["top_score"][live].view(np.uint32))


def _sub(data, lo, hi):
    off = data["offsets"]
    return data["seq_bytes"][off[lo]:off[hi]], off[lo:hi + 1] - off[lo]


def test_c2_index_statistics(c2):
    clf, _ = c2
    st = clf.stats()
    assert st["n_docs"] == 6000 and st["n_terms"] == 4 ** 10
    # every genome contributes len-9 tokens to each of its two strand docs
    from biospectra_b200 import synth
    lengths = synth.genome_lengths(3000, 10_000_000_000, 2)
    assert st["n_positions"] == int(2 * (lengths - 9).sum())
    assert clf.memory_info()["positional_bytes"] > 0          # the positional postings are resident next to the tables
    nt, norm = clf.doc_info(0)
    assert nt == int(lengths[0]) - 9


def test_c2_three_scorers_agree(c2, monkeypatch):
    clf, data = c2
    n = 6000
    seq, off = _sub(data, 0, n)
    base = clf.classify_packed(seq, off)
    routes = clf.last_routes(n)
    assert (base["status"] == 0).all() and (base["n_hits"] >= 1).all()
    assert (routes == 3).all()              # ~1 % of the reads carry an N: their gap clause is evaluated sparsely
    assert (base["top_score"][:, 0] > 0).all()
    monkeypatch.setenv("BSP_DISABLE_GAP", "1")                               # the same reads through the positional scorer
    _same(clf.classify_packed(seq, off), base)
    r2 = clf.last_routes(n)
    assert 0 < (r2 == 2).sum() < 0.05 * n and ((r2 == 2) | (r2 == 3)).all()
    monkeypatch.delenv("BSP_DISABLE_GAP")
    monkeypatch.setenv("BSP_FILTER", "0")                                    # every doc exact, fp64
    _same(clf.classify_packed(seq, off), base)
    assert set(np.unique(clf.last_routes(n))) <= {1, 2}
    monkeypatch.delenv("BSP_FILTER")
    monkeypatch.setenv("BSP_DISABLE_DENSE", "1")                             # positional postings, literal sloppy sweep
    m = 400
    seq2, off2 = _sub(data, 0, m)
    pos = clf.classify_packed(seq2, off2)
    assert (clf.last_routes(m) == 2).all()
    _same(pos, {k: v[:m] for k, v in base.items()})


def test_c2_batches_split_and_entry_points_agree(c2):
    import torch
    clf, data = c2
    n = 40_000
    whole = clf.classify_packed(data["seq_bytes"], data["offsets"])
    for lo, hi in ((0, 1), (1, 17_001), (17_001, 40_000)):
        seq, off = _sub(data, lo, hi)
        _same(clf.classify_packed(seq, off), {k: v[lo:hi] for k, v in whole.items()})
    _same(clf.classify_packed(data["seq_bytes"], data["offsets"]), whole)           # idempotent
    dev = torch.device("cuda", 0)
    o = {k: torch.empty(n, dtype=torch.int32, device=dev) for k in ("status", "label", "n_hits", "n_top")}
    o["top_doc"] = torch.empty(n, 10, dtype=torch.int32, device=dev)
    o["top_score"] = torch.empty(n, 10, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    clf.classify_device(n, data["d_seq"].data_ptr(), data["d_off"].data_ptr(), int(data["offsets"][-1]), o["status"].data_ptr(),
                        o["label"].data_ptr(), o["n_hits"].data_ptr(), o["n_top"].data_ptr(), o["top_doc"].data_ptr(),
                        o["top_score"].data_ptr())
    _same({k: v.cpu().numpy() for k, v in o.items()}, whole)
    # collector order: scores descending, equal scores by ascending doc id
    s, d, nt = whole["top_score"], whole["top_doc"], whole["n_top"]
    for j in range(9):
        live = nt > j + 1
        assert (s[live, j] >= s[live, j + 1]).all()
        tie = live & (s[:, j] == s[:, j + 1])
        assert (d[tie, j] < d[tie, j + 1]).all()


# ------------------------------------------------------------------ unequal document lengths against the oracle
@pytest.mark.parametrize("min_strand", [False, True])
def test_unequal_length_reference_against_oracle(oracle_lib, min_strand, monkeypatch):
    """48 genomes of 0.1 - 1.6 Mbp (norm bytes differ between documents, unlike the equal-length C1 reference), 12,000
    reads at 4 % error + reads with N: the configurations the shipped json does not exercise -- the code default
    kmer_skips = 5 (Configuration.java:37-44), bm25 (Configuration.java:169-179), min_strand_kmer -- each through the
    filter kernel and through the kernel the router picks without it, bit for bit against the oracle."""
    import common
    from biospectra_b200 import synth
    rng = np.random.Generator(np.random.PCG64(77))
    lengths = np.exp(rng.uniform(np.log(1.0e5), np.log(1.6e6), 48)).astype(np.int64)
    genomes = [synth.genome_numpy(7, g, int(lengths[g])) for g in range(len(lengths))]
    reads, _src = synth.sample_reads_numpy(genomes, 12_000, 100, 0.04, 17)
    # every 50th read carries an N in the middle (its window-spanning clauses become gap clauses)
    reads = [bytes(b"N"[0] if j == len(r) // 2 else c for j, c in enumerate(r)) if i % 50 == 0 else r for i, r in enumerate(reads)]
    ents = [g.tobytes().decode() for g in genomes]
    ox = common.build_oracle(oracle_lib, ents, 10, min_strand)
    norms = {ox.doc_info(d)[1] for d in range(ox.stats()["n_docs"])}
    assert len(norms) >= 4                                                  # several norm classes
    g = None
    try:
        for kw, okw in ((dict(), dict()), (dict(kmer_skips=5), dict(skips=5)), (dict(scoring_algorithm="bm25"), dict(scoring=1)),
                        (dict(kmer_skips=5, scoring_algorithm="bm25", query_algorithm="CHAIN_PROXIMITY"), dict(skips=5, scoring=1, algorithm=1))):
            q = dict(dict(skips=0, algorithm=2, scoring=0, msm=0.5, threads=8), **okw)
            ro = ox.classify(reads, **q)
            conf = dict(dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf"), **kw)
            if g is None:
                g = common.build_gpu(ents, 10, min_strand, conf)
                assert g.stats() == ox.stats()
            else:
                from biospectra_b200 import Configuration
                g.set_query(Configuration(index_path="", kmer_size=10, min_strand_kmer=min_strand, full_topn=1, **conf))
            for filt in ("1", "0"):
                monkeypatch.setenv("BSP_FILTER", filt)
                rg = g.classify_batch(reads)
                common.assert_classify_equal(rg, ro, reads, "unequal lengths min_strand=%s %s filter=%s" % (min_strand, kw, filt))
    finally:
        if g is not None:
            g.close()


def test_bm25_filter_with_more_than_16_norm_classes(oracle_lib, monkeypatch):
    """60 genomes of 3 kbp - 800 kbp (a 260 x range of document lengths: 8 norm bytes per factor 4, 25+ classes): BM25 stays on
    the filter kernel (per-class byte LUTs for up to 32 classes; 16 in round 1) and agrees with the oracle bit for bit, also
    for NAIVE_KMER and with the all-exact kernel beside it"""
    import common
    from biospectra_b200 import synth, Configuration
    rng = np.random.Generator(np.random.PCG64(5))
    lengths = np.exp(rng.uniform(np.log(3.0e3), np.log(8.0e5), 60)).astype(np.int64)
    genomes = [synth.genome_numpy(9, g, int(lengths[g])) for g in range(len(lengths))]
    reads, _src = synth.sample_reads_numpy(genomes, 4_000, 100, 0.03, 23)
    ents = [g.tobytes().decode() for g in genomes]
    ox = common.build_oracle(oracle_lib, ents, 8, False)
    norms = {ox.doc_info(d)[1] for d in range(ox.stats()["n_docs"])}
    assert 16 < len(norms) <= 32, len(norms)
    g = common.build_gpu(ents, 8, False, dict(kmer_skips=0, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="bm25"))
    try:
        for alg, a in (("PAIRED_PROXIMITY", 2), ("NAIVE_KMER", 0)):
            ro = ox.classify(reads, skips=0, algorithm=a, scoring=1, msm=0.5, threads=8)
            g.set_query(Configuration(index_path="", kmer_size=8, full_topn=1, kmer_skips=0, query_algorithm=alg, scoring_algorithm="bm25"))
            for filt in ("1", "0"):
                monkeypatch.setenv("BSP_FILTER", filt)
                rg = g.classify_batch(reads)
                common.assert_classify_equal(rg, ro, reads, "bm25, %d norm classes, %s, filter=%s" % (len(norms), alg, filt))
                if filt == "1":
                    assert (g.last_routes(len(reads))[ro["status"] == 0] == 3).mean() > (0.9 if a == 2 else 0.5)   # (NAIVE_KMER: 24 B per clause and class)
    finally:
        g.close()
