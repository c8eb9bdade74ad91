#!/usr/bin/env python
"""Generates tests/golden/lucene_*.json and fasta_reader.json BY EXECUTING THE REFERENCE'S OWN CLASS FILES.

There is no JVM in this image, so `oracle/minijvm.py` (a small bytecode interpreter, test infrastructure) runs the
unmodified bytes of `/root/reference/libs/lucene-core-5.3.1.jar` and `fasta-utils-1.1.jar`:

* lucene_leaves.json  — `util/SmallFloat`, `DefaultSimilarity` (NORM_TABLE <clinit>, computeNorm/lengthNorm/
  encodeNormValue, idf, tf, queryNorm, coord, sloppyFreq), `BM25Similarity` (idf, NORM_TABLE, computeWeight → cache[],
  avgFieldLength), `TFIDFSimilarity$IDFStats#normalize`.
* lucene_sloppy.json  — `search/SloppyPhraseScorer#phraseFreq` (+ PhraseQueue/PhrasePositions/ConjunctionDISI) over
  position lists: every SURVEY App. B row plus seeded random cases, distinct and repeated terms, several slops.
* lucene_search.json  — whole `IndexSearcher.search(BooleanQuery, TopScoreDocCollector.create(10))` runs
  (BooleanWeight/BooleanScorer/MinShouldMatchSumScorer/TermScorer/SloppyPhraseScorer/TFIDF+BM25/HitQueue) on small
  corpora; only the storage under `LeafReader` is a stub (oracle/lucene_exec.py).  Tokens and clauses are fed in as
  `lucene/KmerSequenceTokenizer.java` / `classify/Classifier.java:113-261` produce them (those are Java *sources*,
  restated in oracle/spec_oracle.py; no compiler exists here to run them).
* fasta_reader.json   — `org/yeastrc/fasta/FASTAReader#readNext` on hand-made files (CRLF, ';' lines, lower case,
  interior blanks, errors).

Both oracles (tests/test_lucene_golden.py) and the CUDA path (`-m gpu`) must reproduce these bit for bit.
Run from the repo root (needs /root/reference; takes a few minutes):  python tests/golden/make_lucene_golden.py
"""
import hashlib
import json
import os
import random
import struct
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common            # noqa: E402
import lucene_exec as L  # noqa: E402
import minijvm as J      # noqa: E402
import spec_oracle as so  # noqa: E402

P = L.P
REF = "/root/reference"


def bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def from_bits(b):
    return struct.unpack("<f", struct.pack("<I", b))[0]


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def write(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def executed_summary(vm):
    """Which jar methods really ran (evidence for the reader of the golden file)."""
    return {k: v for k, v in sorted(vm.executed.items()) if k.startswith(("org/apache/lucene", "org/yeastrc"))}


# ------------------------------------------------------------------------------------------------ leaves
def make_leaves():
    lx = L.LuceneExec("tfidf")
    vm = lx.vm
    SF = P + "util/SmallFloat"
    DS = P + "search/similarities/DefaultSimilarity"
    out = {}
    out["byte315ToFloat"] = [bits(vm.call_static(SF, "byte315ToFloat", "(B)F", b - 256 if b > 127 else b)) for b in range(256)]
    rng = random.Random(315)
    fl = [0, 1, 0x7F800000, 0x80000000, 0x3F800000, 0x3F000000, 0x33800000, 0x4F800000, 0x30000000, 0x2FFFFFFF, 0x30200000,
          0x501FFFFF, 0x50000000, 0x4FFFFFFF, 0x7F7FFFFF, 0x00800000, 0xBF800000]
    fl += [rng.getrandbits(32) for _ in range(300)]
    fl += [bits(J.f32(1.0 / (n ** 0.5))) for n in range(1, 400)]
    out["floatToByte315"] = [[b, vm.call_static(SF, "floatToByte315", "(F)B", from_bits(b)) & 0xFF] for b in fl
                             if from_bits(b) == from_bits(b)]
    ds = lx.similarity
    vm.init_class(vm.load(DS))
    out["tfidf_norm_table"] = [bits(x) for x in vm.load(DS).statics["NORM_TABLE"].a]
    lens = list(range(0, 300)) + [1000, 4096, 65536, 1999991, 3299991, 5000000, 6499991, 10000000, 2147483519] + \
        [rng.randint(300, 8000000) for _ in range(200)]
    norms = []
    for n in lens:
        fis = vm.new_init(P + "index/FieldInvertState", "(Ljava/lang/String;)V", "sequence")
        vm.call_virtual(fis, "setLength", "(I)V", n)
        vm.call_virtual(fis, "setBoost", "(F)V", 1.0)
        norms.append([n, vm.call_virtual(ds, "computeNorm", "(L%sindex/FieldInvertState;)J" % P, fis) & 0xFF])
    out["tfidf_computeNorm"] = norms
    grid = [(df, n) for n in (1, 2, 4, 20, 72, 600, 6000, 36000, 1000000) for df in sorted({0, 1, 2, 3, n // 3, n // 2, n - 1, n}) if 0 <= df <= n]
    grid += [(rng.randint(0, 6000), 6000) for _ in range(200)]
    out["tfidf_idf"] = [[df, n, bits(vm.call_virtual(ds, "idf", "(JJ)F", df, n))] for df, n in grid]
    fr = [1, 2, 3, 4, 5, 6, 7, 8, 14, 15, 16, 100, 1000] + [bits(x) for x in ()]
    freqs = [float(x) for x in fr] + [0.5, 1.0 / 3, 1.5, 1 + J.f32(1.0 / 3), J.f32(1.0 / 6), 0.125, 2.5, J.f32(J.f32(1.0 / 3) + J.f32(1.0 / 3))]
    out["tfidf_tf"] = [[bits(J.f32(f)), bits(vm.call_virtual(ds, "tf", "(F)F", J.f32(f)))] for f in freqs]
    sums = [J.f32(x) for x in (0.0, 1.0, 4.523262, 46.0, 91.5, 1e-30, 3e38, 123.456)] + [J.f32(rng.uniform(0.1, 400)) for _ in range(60)]
    out["tfidf_queryNorm"] = [[bits(s), bits(vm.call_virtual(ds, "queryNorm", "(F)F", s))] for s in sums]
    out["tfidf_coord"] = [[i, n, bits(vm.call_virtual(ds, "coord", "(II)F", i, n))] for n in (1, 2, 3, 8, 46, 90, 91, 255, 511, 1000)
                          for i in sorted({0, 1, n // 2, n - 1, n})]
    out["sloppyFreq"] = [[d, bits(vm.call_virtual(ds, "sloppyFreq", "(I)F", d))] for d in range(0, 40)]

    # TFIDFSimilarity#computeWeight + IDFStats#normalize for explicit (df...) clause sets
    def stats_objs(dfs, max_doc, sum_ttf):
        cs = vm.new_init(P + "search/CollectionStatistics", "(Ljava/lang/String;JJJJ)V", "sequence", max_doc, max_doc, sum_ttf, sum_ttf)
        ts = [vm.new_init(P + "search/TermStatistics", "(L%sutil/BytesRef;JJ)V" % P,
                          vm.new_init(P + "util/BytesRef", "([B)V", L.jbytes(b"t%d" % i)), df, df * 2) for i, df in enumerate(dfs)]
        return cs, J.JArray("L%ssearch/TermStatistics;" % P, ts)

    CW = "(FL%ssearch/CollectionStatistics;[L%ssearch/TermStatistics;)L%ssearch/similarities/Similarity$SimWeight;" % (P, P, P)
    norm_cases = []
    for dfs_list, max_doc in (([[17], [20], [5, 5]], 20), ([[5800, 5750], [1, 6000], [3000]], 6000), ([[0], [2, 1]], 4)):
        ws = []
        for dfs in dfs_list:
            cs, ts = stats_objs(dfs, max_doc, 1000)
            ws.append(vm.call_virtual(ds, "computeWeight", CW, 1.0, cs, ts))
        ssq = J.f32(0.0)
        for w in ws:
            ssq = J.f32(ssq + vm.call_virtual(w, "getValueForNormalization", "()F"))
        qn = vm.call_virtual(ds, "queryNorm", "(F)F", ssq)
        vals = []
        for w in ws:
            vm.call_virtual(w, "normalize", "(FF)V", qn, 1.0)
            vals.append(bits(w.f[P + "search/similarities/TFIDFSimilarity$IDFStats.value"]))
        norm_cases.append({"dfs": dfs_list, "max_doc": max_doc, "sum_sq": bits(ssq), "query_norm": bits(qn), "values": vals})
    out["tfidf_normalize"] = norm_cases

    # BM25
    lb = L.LuceneExec("bm25")
    vb = lb.vm
    bs = lb.similarity
    B = P + "search/similarities/BM25Similarity"
    vb.init_class(vb.load(B))
    out["bm25_norm_table"] = [bits(x) for x in vb.load(B).statics["NORM_TABLE"].a]
    out["bm25_idf"] = [[df, n, bits(vb.call_virtual(bs, "idf", "(JJ)F", df, n))] for df, n in grid]
    out["bm25_computeNorm"] = []
    for n in lens[:320]:
        fis = vb.new_init(P + "index/FieldInvertState", "(Ljava/lang/String;)V", "sequence")
        vb.call_virtual(fis, "setLength", "(I)V", n)
        vb.call_virtual(fis, "setBoost", "(F)V", 1.0)
        out["bm25_computeNorm"].append([n, vb.call_virtual(bs, "computeNorm", "(L%sindex/FieldInvertState;)J" % P, fis) & 0xFF])
    caches = []
    for max_doc, sum_ttf, dfs in ((20, 39999820, [17]), (6000, 19999943008, [5750, 5800]), (4, 32, [2]), (7, 0, [1]), (3, 1000, [0])):
        cs = vb.new_init(P + "search/CollectionStatistics", "(Ljava/lang/String;JJJJ)V", "sequence", max_doc, max_doc, sum_ttf, sum_ttf)
        ts = J.JArray("L%ssearch/TermStatistics;" % P, [vb.new_init(P + "search/TermStatistics", "(L%sutil/BytesRef;JJ)V" % P,
                      vb.new_init(P + "util/BytesRef", "([B)V", L.jbytes(b"t%d" % i)), df, df * 2) for i, df in enumerate(dfs)])
        avg = vb.call_virtual(bs, "avgFieldLength", "(L%ssearch/CollectionStatistics;)F" % P, cs)
        w = vb.call_virtual(bs, "computeWeight", CW, 1.0, cs, ts)
        vb.call_virtual(w, "normalize", "(FF)V", 1.0, 1.0)
        st = w.f
        caches.append({"max_doc": max_doc, "sum_ttf": sum_ttf, "dfs": dfs, "avgdl": bits(avg),
                       "cache": [bits(x) for x in st[B + "$BM25Stats.cache"].a],
                       "weight": bits(st[B + "$BM25Stats.weight"])})
    out["bm25_weight_cache"] = caches
    out["executed"] = executed_summary(vm)
    out["executed_bm25"] = executed_summary(vb)
    out["math_log_calls"] = len(vm.math_log_calls) + len(vb.math_log_calls)
    return out


# ------------------------------------------------------------------------------------------------ sloppy
APP_B = [([5], [6], 2), ([5], [7], 2), ([5], [8], 2), ([5], [9], 2), ([5], [4], 2), ([5], [3], 2), ([5, 100], [6, 101], 2),
         ([5, 100], [6, 103], 2), ([5, 6, 7], [8], 2), ([3, 50], [20, 51, 52], 2), ([10], [16], 7), ([10], [18], 7), ([10], [19], 7)]
APP_B_SAME = [([5, 6], 2), ([5, 6, 7], 2), ([5, 9], 2), ([5], 2)]


def make_sloppy():
    lx = L.LuceneExec("tfidf")
    cases = []

    def run(p0, p1, slop, same):
        r = lx.phrase_freq(p0, p1, slop, same)
        cases.append({"p0": list(p0), "p1": None if same else list(p1), "slop": slop, "same": same,
                      "freq_bits": bits(0.0 if r is None else r)})

    for p0, p1, s in APP_B:
        run(p0, p1, s, False)
    for p, s in APP_B_SAME:
        run(p, None, s, True)
    rng = random.Random(77)
    for _ in range(500):
        span = rng.choice((8, 12, 20, 40))
        pts = rng.sample(range(span), rng.randint(2, min(span, 10)))
        cut = rng.randint(1, len(pts) - 1)
        run(sorted(pts[:cut]), sorted(pts[cut:]), rng.choice((1, 2, 2, 2, 3, 7, 7, 8, 13)), False)
    for _ in range(300):
        span = rng.choice((6, 10, 16, 30))
        pts = sorted(rng.sample(range(span), rng.randint(1, min(span, 9))))
        run(pts, None, rng.choice((1, 2, 2, 2, 3, 7, 7, 8)), True)
    for run_len in range(1, 12):                      # homopolymer runs: consecutive positions of one term
        for s in (2, 7):
            run(list(range(3, 3 + run_len)), None, s, True)
    return {"cases": cases, "executed": executed_summary(lx.vm)}


# ------------------------------------------------------------------------------------------------ search
def kmer_of(t, k):
    return "".join("ACGT"[(t >> (2 * (k - 1 - i))) & 3] for i in range(k))


def lucene_docs(ix):
    """so.Index → per-doc token lists of Lucene term bytes (Base64 of the packed k-mer, SequenceCompressFilter.java:70-86)."""
    docs = [[None] * d.ntok for d in ix.docs]
    for t, pl in ix.postings.items():
        tb = so.term_text(kmer_of(t, ix.k)).encode()
        for d, poss in pl.items():
            for p in poss:
                docs[d][p] = tb
    assert all(x is not None for d in docs for x in d)
    return docs


def lucene_clauses(cl, k):
    out = []
    for c in cl:
        if c[0] == "T":
            out.append(("term", so.term_text(kmer_of(c[1], k)).encode()))
        else:
            out.append(("phrase", so.term_text(kmer_of(c[1], k)).encode(), so.term_text(kmer_of(c[2], k)).encode(), c[3]))
    return out


def ref_reads(path, n, start=0, max_len=None):
    ents = so.read_fasta(open(path, "rb").read().decode("ascii", "replace"))
    seqs = [s for _h, s in ents][start:start + n]
    return [s[:max_len] if max_len else s for s in seqs]


def plant(rng, reads, n_entries, flank=(20, 200), mut=0.03):
    """Synthetic reference entries that contain (mutated) copies of the reads, so that reference-held queries hit."""
    ents = []
    per = max(1, (len(reads) + n_entries - 1) // n_entries)
    for e in range(n_entries):
        parts = [common.rand_seq(rng, rng.randint(*flank))]
        for r in reads[e * per:(e + 1) * per] + rng.sample(reads, min(2, len(reads))):
            s = "".join(c if c in "ACGT" else rng.choice("ACGT") for c in r)
            if rng.random() < 0.5:
                s = so.reverse_complement(s)
            s = "".join(rng.choice("ACGT") if rng.random() < mut else c for c in s)
            parts += [s, common.rand_seq(rng, rng.randint(*flank))]
        ents.append("".join(parts))
    return ents


def search_cases():
    bn = os.path.join(REF, "benchmark")
    sim = lambda e: os.path.join(bn, "bacteria_ncbi_simulated", "sample_%s.fa" % e)  # noqa: E731
    shc = sorted(os.listdir(os.path.join(bn, "simHC.20.500")))
    cases = []
    seed = 0
    for k, ms, skips, alg, scoring, msm, junk in [
        (3, False, 0, "NAIVE_KMER", "tfidf", 0.5, 0.0), (3, False, 0, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.0),
        (3, True, 0, "CHAIN_PROXIMITY", "bm25", 0.5, 0.02), (4, True, 1, "PAIRED_PROXIMITY", "tfidf", 0.3, 0.0),
        (5, False, 0, "PAIRED_PROXIMITY", "bm25", 0.5, 0.02), (5, False, 2, "CHAIN_PROXIMITY", "tfidf", 0.0, 0.0),
        (6, True, 0, "NAIVE_KMER", "bm25", 0.9, 0.03), (8, False, 5, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.01),
        (10, False, 0, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.0), (10, False, 5, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.0),
        (10, True, 0, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.01), (10, False, 0, "NAIVE_KMER", "tfidf", 0.5, 0.0),
        (10, False, 0, "CHAIN_PROXIMITY", "tfidf", 0.5, 0.0), (10, False, 0, "PAIRED_PROXIMITY", "bm25", 0.5, 0.0),
        (10, True, 5, "PAIRED_PROXIMITY", "bm25", 1.0, 0.0), (12, False, 1, "CHAIN_PROXIMITY", "tfidf", 0.0, 0.0),
        (13, False, 0, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.0), (16, True, 0, "PAIRED_PROXIMITY", "tfidf", 0.5, 0.01),
        (20, False, 5, "NAIVE_KMER", "bm25", 0.5, 0.0), (24, False, 0, "CHAIN_PROXIMITY", "tfidf", 0.5, 0.0),
    ]:
        seed += 1
        ents = common.make_corpus(900 + seed, n_entries=(4, 8), length=(20, 300), junk=junk)
        ents = [so.java_trim(x.upper()) or "ACGT" for x in ents]
        reads = common.make_reads(1900 + seed, ents, n=16, max_len=120)
        cases.append(dict(name="adv%02d" % seed, k=k, min_strand=ms, skips=skips, algorithm=alg, scoring=scoring, msm=msm,
                          entries=ents, reads=reads))
    # reference-held queries: simulated 100 bp reads at 0 / 4 / 19 % error, and simHC Sanger reads with X/N masks
    rng = random.Random(4242)
    for tag, path, n, start, k, skips, alg, scoring, ms in [
        ("sim0.0", sim("0.0"), 24, 0, 10, 0, "PAIRED_PROXIMITY", "tfidf", False),
        ("sim0.04", sim("0.04"), 24, 100, 10, 0, "PAIRED_PROXIMITY", "tfidf", False),
        ("sim0.04_skips5", sim("0.04"), 24, 200, 10, 5, "PAIRED_PROXIMITY", "tfidf", False),
        ("sim0.04_naive", sim("0.04"), 16, 300, 10, 0, "NAIVE_KMER", "tfidf", False),
        ("sim0.04_chain_bm25", sim("0.04"), 16, 400, 10, 0, "CHAIN_PROXIMITY", "bm25", False),
        ("sim0.19_minstrand", sim("0.19"), 16, 0, 10, 0, "PAIRED_PROXIMITY", "tfidf", True),
        ("sim0.08_k8", sim("0.08"), 16, 0, 8, 0, "PAIRED_PROXIMITY", "bm25", False),
        ("simHC_a", os.path.join(bn, "simHC.20.500", shc[0]), 6, 0, 10, 0, "PAIRED_PROXIMITY", "tfidf", False),
        ("simHC_b_skips5", os.path.join(bn, "simHC.20.500", shc[7]), 6, 10, 10, 5, "PAIRED_PROXIMITY", "tfidf", False),
        ("simHC_c_naive_bm25", os.path.join(bn, "simHC.20.500", shc[13]), 4, 20, 10, 0, "NAIVE_KMER", "bm25", False),
        ("simHC_d_chain", os.path.join(bn, "simHC.20.500", shc[19]), 4, 30, 10, 0, "CHAIN_PROXIMITY", "tfidf", True),
    ]:
        reads = ref_reads(path, n, start)
        ents = plant(rng, reads, 6)
        cases.append(dict(name=tag, source=os.path.relpath(path, REF), k=k, min_strand=ms, skips=skips, algorithm=alg,
                          scoring=scoring, msm=0.5, entries=ents, reads=reads))
    # > 2048 documents: BooleanScorer's 2,048-doc windows, > 10-way ties, top-10 cut among equal scores
    rng = random.Random(99)
    base = common.rand_seq(rng, 60)
    ents = []
    for i in range(1100):
        a = rng.randint(0, 40)
        s = base[a:a + rng.randint(8, 20)] if rng.random() < 0.7 else common.rand_seq(rng, rng.randint(6, 24))
        ents.append(s)
    reads = [base[5:45], base[20:60], base[0:30], common.rand_seq(rng, 40), base[10:22] + "N" + base[23:50]]
    for alg, scoring in (("PAIRED_PROXIMITY", "tfidf"), ("NAIVE_KMER", "bm25")):
        cases.append(dict(name="wide_%s_%s" % (alg[:5].lower(), scoring), k=4, min_strand=False, skips=0, algorithm=alg, scoring=scoring,
                          msm=0.5, entries=ents, reads=reads))
    # BooleanQuery.setMaxClauseCount(10000) (Classifier.java:110): 10,000 clauses run, 10,001 throw TooManyClauses
    rng = random.Random(10000)
    cases.append(dict(name="maxclause", k=3, min_strand=False, skips=0, algorithm="NAIVE_KMER", scoring="tfidf", msm=0.5,
                      entries=[common.rand_seq(rng, 90), common.rand_seq(rng, 40)],
                      reads=[common.rand_seq(rng, 10002), common.rand_seq(rng, 10003)]))
    return cases


def make_search():
    out = []
    executed = {}
    t0 = time.time()
    for case in search_cases():
        k, ms = case["k"], case["min_strand"]
        ix = so.Index(k, ms)
        for i, s in enumerate(case["entries"]):
            ix.add_entry("f%d.fa" % i, "h%d" % i, "", s)
        lx = L.LuceneExec(case["scoring"])
        lx.index(lucene_docs(ix))
        alg = so.ALGORITHMS[case["algorithm"]]
        results = []
        for q in case["reads"]:
            toks = so.tokenize(q, k, case["skips"])
            cl = so.build_clauses(toks, alg, ms)
            if cl is None:
                results.append(None)          # Classifier.java:223-227: null query → NPE → read dropped
                continue
            try:
                query, msm = lx.build_query(lucene_clauses(cl, k), case["msm"])
            except J.JavaThrow as t:      # BooleanQuery$TooManyClauses (> 10,000): the reference logs and skips the read
                assert t.obj.cls.name.endswith("TooManyClauses"), t
                results.append(None)
                continue
            w = lx.weights(query)
            r = lx.search(query, 10)
            results.append({"n_clauses": len(cl), "msm": msm, "weights": [bits(x) for x in w], "total_hits": r["total_hits"],
                            "max_score": None if r["max_score"] != r["max_score"] else bits(r["max_score"]),
                            "top": [[d, bits(s)] for d, s in r["hits"]]})
        for kk, v in executed_summary(lx.vm).items():
            executed[kk] = executed.get(kk, 0) + v
        rec = dict(case)
        rec["norm_bytes"] = lx.norm_bytes()
        rec["n_docs"] = len(ix.docs)
        rec["results"] = results
        out.append(rec)
        print("  %-22s docs %5d reads %3d  (%.0f s)" % (case["name"], len(ix.docs), len(case["reads"]), time.time() - t0), flush=True)
    return {"cases": out, "executed": executed}


# ------------------------------------------------------------------------------------------------ FASTA
FASTA_FILES = {
    "plain": ">e1 first\nACGT\nacgtnn\n>e2\nGGGG\n",
    "crlf": ">e1\r\nACGT\r\nTTTT\r\n>e2 x y\r\nCC CC\r\n",
    "cr_only": ">e1\rACGT\rGG\r",
    "comments": ";lead comment\n>e1\n;inner comment\nAC\n;again\nGT\n>e2\nTT\n",
    "blank_tail": ">e1\nACGT  \n  \n\n>e2\n  AC GT\t\n",
    "lower_iupac": ">mixed|case\nacgtRYKMnx-*\n",
    "no_trailing_newline": ">e1\nACGT",
    "ctrl_a": ">gi|1|first\x01gi|2|second\nACGT\n",
    "header_only_then_entry": ">e1\n>e2\nAC\n",
    "header_at_eof": ">e1\nAC\n>e2\n",
    "no_header": "ACGT\n>e1\nAC\n",
    "empty": "",
    "header_leading_blank": "> e1 spaced\nACGT\n",
    "header_empty": ">\nACGT\n",
    "header_tab": ">e1\tdesc with  two spaces \nAC\n",
    "second_header_bad": ">ok\nAC\n> bad\nGT\n",
    "semicolon_first_after_header": ">e1\n;c\n>e2\nAC\n",
    "only_comment": ";just a comment\n",
    "blank_first_line": "\n>e1\nAC\n",
    "interior_gt": ">e1\nAC>GT\nTT\n",
    "high_bytes": ">e1 caf\xe9\nAC\xe9GT\n",
}


def make_fasta():
    vm = J.VM([os.path.join(L.REF_LIBS, "fasta-utils-1.1.jar")])
    out = []
    with tempfile.TemporaryDirectory() as td:
        for name, text in FASTA_FILES.items():
            path = os.path.join(td, name + ".fa")
            with open(path, "wb") as f:
                f.write(text.encode("latin-1"))
            rec = {"name": name, "text": text, "entries": [], "error": None}
            try:
                rd = vm.call_static("org/yeastrc/fasta/FASTAReader", "getInstance", "(Ljava/lang/String;)Lorg/yeastrc/fasta/FASTAReader;", path)
                while True:
                    e = vm.call_virtual(rd, "readNext", "()Lorg/yeastrc/fasta/FASTAEntry;")
                    if e is None:
                        break
                    rec["entries"].append([vm.call_virtual(e, "getHeaderLine", "()Ljava/lang/String;"),
                                           vm.call_virtual(e, "getSequence", "()Ljava/lang/String;")])
            except J.JavaThrow as t:
                rec["error"] = t.obj.cls.name
            out.append(rec)
    return {"cases": out, "executed": executed_summary(vm)}


def main():
    which = set(sys.argv[1:]) or {"leaves", "sloppy", "search", "fasta"}
    meta = {"generator": "tests/golden/make_lucene_golden.py: the reference's class files executed by oracle/minijvm.py",
            "lucene_jar_sha256": sha256(L.LUCENE_JAR), "fasta_jar_sha256": sha256(os.path.join(L.REF_LIBS, "fasta-utils-1.1.jar"))}
    if "leaves" in which:
        write("lucene_leaves.json", dict(meta, **make_leaves()))
    if "sloppy" in which:
        write("lucene_sloppy.json", dict(meta, **make_sloppy()))
    if "fasta" in which:
        write("fasta_reader.json", dict(meta, **make_fasta()))
    if "search" in which:
        write("lucene_search.json", dict(meta, **make_search()))


if __name__ == "__main__":
    main()
