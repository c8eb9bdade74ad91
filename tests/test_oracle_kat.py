"""Known-answer tests that pin BOTH oracles (literal Python spec oracle and the fast C oracle) to the
hand-traced values of SURVEY.md Appendix B.  The reference ships no tests or golden outputs and cannot
run here (no JVM), so these values -- derived from the reference sources and the Lucene 5.3.1 class
files -- are what "parity" is anchored on."""
import struct

import numpy as np
import pytest

import spec_oracle as so


def f32bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


# ------------------------------------------------------------------ term encoding (A.2)
@pytest.mark.parametrize("kmer,term", [
    ("ACGTACGTAC", "GxsQ"), ("AAAAAAAAAA", "AAAA"), ("TTTTTTTTTT", "///w"), ("GATTACAGAT", "jxIw"),
    ("ACGT", "Gw=="), ("ACGTA", "GwA="), ("ACGT" * 5, "GxsbGxs="),
])
def test_term_text(kmer, term):
    # lucene/SequenceCompressFilter.java:70-86 + utils/SequenceHelper.java:225-271
    assert so.term_text(kmer) == term


def test_compress_bytes():
    assert so.compress("ACGTACGTAC") == bytes([0x1B, 0x1B, 0x10])


def test_min_strand_canonical(oracle_lib):
    assert so.canonical("TTTTTTTTTT", True) == "AAAAAAAAAA"
    assert so.canonical("ACGTACGTAC", True) == "ACGTACGTAC"        # revcomp GTACGTACGT is larger
    keys, _ = oracle_lib.tokenize(b"TTTTTTTTTT", 10, 0, True)
    assert keys.tolist() == [0]
    keys, _ = oracle_lib.tokenize(b"ACGTACGTAC", 10, 0, True)
    assert keys.tolist() == [so.pack_kmer("ACGTACGTAC")]


# ------------------------------------------------------------------ SmallFloat / norms / idf (A.5, A.6)
def test_small_float(oracle_lib):
    L = oracle_lib.lib()
    for impl_f2b, impl_b2f in ((so.float_to_byte315, so.byte315_to_float),
                               (lambda x: L.ora_float_to_byte315(float(x)), lambda b: L.ora_byte315_to_float(b))):
        assert impl_f2b(np.float32(1.0)) == 124
        assert impl_f2b(np.float32(0.5)) == 120
        assert float(impl_b2f(124)) == 1.0
        assert f32bits(impl_b2f(1)) == f32bits(np.float32(5.8207661e-10))
        assert f32bits(impl_b2f(255)) == f32bits(np.float32(7.5161928e9))
        assert float(impl_b2f(0)) == 0.0
    for b in range(256):       # both implementations agree on the whole table and round-trip
        assert f32bits(so.byte315_to_float(b)) == f32bits(L.ora_byte315_to_float(b))
        if b:
            assert so.float_to_byte315(so.byte315_to_float(b)) == b


@pytest.mark.parametrize("n,byte,decoded", [
    (1, 124, 1.0), (2, 121, 0.625), (10, 117, 0.3125), (91, 110, 0.09375), (100, 110, 0.09375), (1000, 104, 0.03125),
    (1_999_991, 81, 6.1035156e-4), (3_299_991, 80, 4.8828125e-4), (5_000_000, 79, None), (10_000_000, 77, None),
])
def test_norm_byte(n, byte, decoded, oracle_lib):
    assert so.norm_byte(n) == byte
    assert oracle_lib.lib().ora_norm_byte(n) == byte
    if decoded is not None:
        assert f32bits(so.byte315_to_float(byte)) == f32bits(np.float32(decoded))


@pytest.mark.parametrize("df,n,idf", [(17, 20, 1.1053605), (20, 20, 0.95120984), (5800, 6000, 1.0337292), (1, 6000, 9.0063677)])
def test_idf_tfidf(df, n, idf, oracle_lib):
    assert f32bits(so.idf_tfidf(df, n)) == f32bits(np.float32(idf))
    assert f32bits(oracle_lib.lib().ora_idf_tfidf(df, n)) == f32bits(np.float32(idf))


def test_idf_bm25_matches(oracle_lib):
    for df, n in ((0, 4), (1, 4), (4, 4), (17, 20), (5800, 6000)):
        assert f32bits(so.idf_bm25(df, n)) == f32bits(oracle_lib.lib().ora_idf_bm25(df, n))


# ------------------------------------------------------------------ tokenizer (A.2)
def _toks_spec(seq, k, skips):
    return [(w, o) for w, o in so.tokenize(seq, k, skips)]


def _toks_c(ol, seq, k, skips):
    keys, offs = ol.tokenize(seq.encode(), k, skips)
    return [(int(a), int(b)) for a, b in zip(keys, offs)]


def test_tokenizer_n_handling(oracle_lib):
    # the emit that follows a dropped window advances by 1 only (KmerSequenceTokenizer.java:119-153)
    exp = [("ACG", 0), ("ACG", 5), ("CGT", 6)]
    assert _toks_spec("ACGTNACGTA", 3, 1) == exp
    assert _toks_c(oracle_lib, "ACGTNACGTA", 3, 1) == [(so.pack_kmer(w), o) for w, o in exp]


def test_tokenizer_stride(oracle_lib):
    exp = [("ACG", 0), ("GTA", 2), ("ACG", 4), ("GTA", 6)]
    assert _toks_spec("ACGTACGTAC", 3, 1) == exp
    assert _toks_c(oracle_lib, "ACGTACGTAC", 3, 1) == [(so.pack_kmer(w), o) for w, o in exp]
    read = ("ACGT" * 25)
    assert [o for _, o in _toks_spec(read, 10, 5)] == list(range(0, 91, 6))
    assert [o for _, o in _toks_c(oracle_lib, read, 10, 5)] == list(range(0, 91, 6))
    assert len(_toks_spec(read, 10, 0)) == 91


def test_tokenizer_rejects_lowercase_and_iupac(oracle_lib):
    for seq in ("acgtaCGTAC", "ACGTRCGTAC", "ACGT CGTAC", "ACGTXCGTAC", "ACGTNCGTAC"):
        assert _toks_spec(seq, 5, 0) == [(seq[5:10], 5)]
        assert _toks_c(oracle_lib, seq, 5, 0) == [(so.pack_kmer(seq[5:10]), 5)]
    assert _toks_spec("acgtacgtac", 5, 0) == [] and _toks_c(oracle_lib, "acgtacgtac", 5, 0) == []
    assert _toks_spec("ACG", 5, 0) == [] and _toks_c(oracle_lib, "ACG", 5, 0) == []
    assert _toks_spec("", 5, 0) == [] and _toks_c(oracle_lib, "", 5, 0) == []


@pytest.mark.parametrize("n_tok,alg,nc,msm", [(91, 0, 91, 45), (91, 1, 90, 45), (91, 2, 46, 23), (16, 2, 8, 4)])
def test_clause_counts(n_tok, alg, nc, msm, oracle_lib):
    # Classifier.java:113-200,257
    toks = [("A" * 10, i) for i in range(n_tok)]
    cl = so.build_clauses(toks, alg, False)
    assert len(cl) == nc and int(0.5 * len(cl)) == msm
    if alg == 2 and n_tok % 2:
        assert cl[-1][0] == "T"
    ox = oracle_lib.OracleIndex(10, False)
    ox.add_entry(b"ACGT" * 30)
    ox.finish()
    skips = 0 if n_tok == 91 else 5
    r = ox.classify([b"ACGT" * 25], skips=skips, algorithm=alg, scoring=0, msm=0.5)
    assert r["n_clauses"][0] == nc and r["msm"][0] == msm
    if skips == 5:
        assert cl[0][3] == 2 and so.build_clauses([("A" * 10, 6 * i) for i in range(16)], 2, False)[0][3] == 7


def test_too_few_tokens_dropped(oracle_lib):
    ox = oracle_lib.OracleIndex(10, False)
    ox.add_entry(b"ACGT" * 30)
    ox.finish()
    r = ox.classify([b"ACGTACGTAC", b"ACGTACGTA", b"N" * 50, b"ACGTACGTACG"], skips=0, algorithm=2, scoring=0, msm=0.5)
    assert r["status"].tolist() == [1, 1, 1, 0]          # 1, 0, 0 and 2 tokens


# ------------------------------------------------------------------ sloppy phrase frequency (A.7)
SLOPPY = [
    # P0, P1, slop, same, expected (as a sum of float32 terms in accumulation order)
    ([5], [6], 2, False, [1.0]), ([5], [7], 2, False, [0.5]), ([5], [8], 2, False, [1 / 3]), ([5], [9], 2, False, []),
    ([5], [4], 2, False, [1 / 3]), ([5], [3], 2, False, []),
    ([5, 100], [6, 101], 2, False, [1.0, 1.0]), ([5, 100], [6, 103], 2, False, [1.0, 1 / 3]),
    ([5, 6, 7], [8], 2, False, [1.0]), ([3, 50], [20, 51, 52], 2, False, [1.0, 0.5]),
    ([10], [16], 7, False, [1 / 6]), ([10], [18], 7, False, [1 / 8]), ([10], [19], 7, False, []),
    ([5, 6], [5, 6], 2, True, [1.0]), ([5, 6, 7], [5, 6, 7], 2, True, [1.0, 1.0]), ([5, 9], [5, 9], 2, True, []),
    ([5], [5], 2, True, []),
]


@pytest.mark.parametrize("p0,p1,slop,same,terms", SLOPPY)
def test_sloppy_freq(p0, p1, slop, same, terms, oracle_lib):
    exp = np.float32(0.0)
    for t in terms:
        exp = np.float32(exp + np.float32(np.float32(1.0) / np.float32(round(1 / t))))
    got_s = so.sloppy_freq(p0, p1, slop, same)
    got_c = oracle_lib.sloppy_freq(p0, p1, slop, same)
    assert f32bits(got_s) == f32bits(exp), (got_s, exp)
    assert f32bits(got_c) == f32bits(exp), (got_c, exp)


# ------------------------------------------------------------------ end-to-end worked example (B.2)
def test_worked_example(oracle_lib):
    ix = so.Index(3, False)
    ix.add_entry("a.fa", "e0", "", "ACGTACGTAC")
    ix.add_entry("a.fa", "e1", "", "ACGTTTTTTT")
    assert ix.max_doc == 4 and [d.ntok for d in ix.docs] == [8] * 4 and [d.norm for d in ix.docs] == [117] * 4
    assert [d.direction for d in ix.docs] == ["forward", "reverse", "forward", "reverse"]
    hits, nc, msm = so.search(ix, "ACGTAC", skips=0, algorithm=0, scoring=0, msm_ratio=0.5)
    assert (nc, msm) == (4, 2)
    assert [(d, f32bits(s)) for d, s in hits] == [(0, 0x3F709E98), (1, 0x3F709E98), (2, 0x3DB59BB7), (3, 0x3DB59BB7)]
    res = so.classify(ix, ">q", "ACGTAC", skips=0, algorithm=0, scoring=0, msm_ratio=0.5)
    assert res["type"] == "VAGUE" and [e["docid"] for e in res["result"]] == [0, 1]
    # C oracle: same bits
    ox = oracle_lib.OracleIndex(3, False)
    ox.add_entry(b"ACGTACGTAC")
    ox.add_entry(b"ACGTTTTTTT")
    ox.finish()
    r = ox.classify([b"ACGTAC"], skips=0, algorithm=0, scoring=0, msm=0.5)
    assert r["n_top"][0] == 4 and r["n_hits"][0] == 2 and r["label"][0] == 1
    assert r["top_doc"][0, :4].tolist() == [0, 1, 2, 3]
    assert [f32bits(x) for x in r["top_score"][0, :4]] == [0x3F709E98, 0x3F709E98, 0x3DB59BB7, 0x3DB59BB7]


def test_worked_example_with_taxonomy():
    lineage = ('[{"taxid":11,"name":"strain x","parent":10,"rank":"no rank"},'
               '{"taxid":10,"name":"Escherichia coli","parent":9,"rank":"species"},'
               '{"taxid":9,"name":"Escherichia","parent":1,"rank":"genus"}]')
    ix = so.Index(3, False)
    ix.add_entry("a.fa", "e0", lineage, "ACGTACGTAC")
    ix.add_entry("a.fa", "e1", "", "ACGTTTTTTT")
    res = so.classify(ix, ">q", "ACGTAC", skips=0, algorithm=0, scoring=0, msm_ratio=0.5)
    # both strands of an entry carry the same lineage (Indexer.java:207): LCA = lowest ranked taxon
    assert (res["type"], res["taxon_rank"], res["taxon_name"]) == ("CLASSIFIED", "species", "Escherichia coli")


def test_reverse_doc_lost_on_out_of_range_char(oracle_lib):
    # utils/SequenceHelper.java:33-35: chars outside 'A'..'Z' throw in getReverseComplement, the
    # exception is swallowed at Indexer.java:227-229 -> only the forward doc exists
    ix = so.Index(3, False)
    ix.add_entry("a.fa", "e0", "", "ACGTAC-GTAC")
    ix.add_entry("a.fa", "e1", "", "ACGTNACGT")
    assert [d.direction for d in ix.docs] == ["forward", "forward", "reverse"]
    ox = oracle_lib.OracleIndex(3, False)
    ox.add_entry(b"ACGTAC-GTAC")
    ox.add_entry(b"ACGTNACGT")
    ox.finish()
    assert ox.stats()["n_docs"] == 3
