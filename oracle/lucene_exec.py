"""Drive the reference's OWN Lucene 5.3.1 class files (`/root/reference/libs/lucene-core-5.3.1.jar`) through
minijvm — TEST INFRASTRUCTURE ONLY, used by `tests/golden/make_lucene_golden.py` to produce the committed
`tests/golden/lucene_*.json` vectors.  Needs `/root/reference` (this container only, never the GPU box).

What runs from the jar, unmodified: `IndexSearcher.search(Query, Collector)` → `createNormalizedWeight`
(`BooleanQuery#rewrite`, `BooleanWeight`, `TermQuery$TermWeight`, `PhraseQuery$PhraseWeight`, `TermContext#build`,
`TFIDFSimilarity#computeWeight/$IDFStats#normalize` or `BM25Similarity`), `BooleanWeight#bulkScorer` →
`BooleanScorer` / `Weight$DefaultBulkScorer` / `MinShouldMatchSumScorer`, `TermScorer`, `SloppyPhraseScorer`
(+ `PhraseQueue`, `PhrasePositions`, `ConjunctionDISI`), `TopScoreDocCollector`, `HitQueue`, `SmallFloat`,
`Similarity#computeNorm`.  What is stubbed here (Python classes extending the jar's abstract classes): only the
storage — a `LeafReader` over in-memory postings (`Fields`/`Terms`/`TermsEnum`/`PostingsEnum`/`NumericDocValues`),
i.e. what `IndexWriter` + the Lucene50 codec would have written to disk and read back.  Positions are token
ordinals (posIncr 1), the norm of a doc is `Similarity#computeNorm(FieldInvertState)` executed from the jar.

The query is assembled with exactly the calls `classify/Classifier.java:113-261` makes
(`new Term(field, BytesRef)`, `new TermQuery`, `PhraseQuery.Builder#setSlop/#add(Term)`, two
`BooleanQuery.Builder`s with `setDisableCoord(false)` and `setMinimumNumberShouldMatch`).
"""
from __future__ import annotations

import os

import minijvm as J

REF_LIBS = "/root/reference/libs"
LUCENE_JAR = os.path.join(REF_LIBS, "lucene-core-5.3.1.jar")

P = "org/apache/lucene/"
NO_MORE_DOCS = 2147483647


def jbytes(b: bytes) -> J.JArray:
    return J.JArray("B", [x - 256 if x > 127 else x for x in b])


class LuceneExec:
    def __init__(self, scoring: str = "tfidf", field: str = "sequence", assertions: bool = False):
        self.vm = vm = J.VM([LUCENE_JAR])
        vm.assertions = assertions
        self.field = field
        sim_cls = P + ("search/similarities/BM25Similarity" if scoring == "bm25" else "search/similarities/DefaultSimilarity")
        self.similarity = vm.new_init(sim_cls, "()V")
        self.scoring = scoring
        self._replace_memory_accounting()
        vm.call_static(P + "search/BooleanQuery", "setMaxClauseCount", "(I)V", 10000)   # Classifier.java:110
        self._define_stubs()
        self.reader = None
        self.searcher = None

    # ------------------------------------------------------------------ the ONLY jar methods not executed
    def _replace_memory_accounting(self):
        """`util/RamUsageEstimator` probes the JVM through reflection / JMX (sun.misc.Unsafe, HotSpotDiagnosticMXBean)
        to size objects for cache accounting.  It feeds no score; its static initialiser and `shallowSizeOfInstance`
        are replaced by the constants a 64-bit HotSpot with compressed oops reports."""
        vm = self.vm
        RUE = P + "util/RamUsageEstimator"

        @vm.native(RUE, "<clinit>")
        def _(vm, a):
            c = vm.load(RUE)
            c.statics.update({"NUM_BYTES_BOOLEAN": 1, "NUM_BYTES_BYTE": 1, "NUM_BYTES_CHAR": 2, "NUM_BYTES_SHORT": 2,
                              "NUM_BYTES_INT": 4, "NUM_BYTES_FLOAT": 4, "NUM_BYTES_LONG": 8, "NUM_BYTES_DOUBLE": 8,
                              "COMPRESSED_REFS_ENABLED": 1, "NUM_BYTES_OBJECT_REF": 4, "NUM_BYTES_OBJECT_HEADER": 12,
                              "NUM_BYTES_ARRAY_HEADER": 16, "NUM_BYTES_OBJECT_ALIGNMENT": 8, "ONE_KB": 1024,
                              "ONE_MB": 1 << 20, "ONE_GB": 1 << 30, "JVM_IS_HOTSPOT_64BIT": 1})

        @vm.native(RUE, "shallowSizeOfInstance")
        def _(vm, a):
            return 32

        @vm.native(RUE, "shallowSizeOf")
        def _(vm, a):
            return 32

    # ------------------------------------------------------------------ storage stubs
    def _define_stubs(self):
        vm = self.vm
        me = self

        def bref_bytes(br) -> bytes:
            arr = br.f[P + "util/BytesRef.bytes"].a
            off = br.f[P + "util/BytesRef.offset"]
            ln = br.f[P + "util/BytesRef.length"]
            return bytes((x & 0xFF) for x in arr[off:off + ln])

        # PostingsEnum over one term's [(doc, [positions...])]
        def pe_next(vm, a):
            p = a[0].p
            p["i"] += 1
            p["pi"] = 0
            return p["list"][p["i"]][0] if p["i"] < len(p["list"]) else NO_MORE_DOCS

        def pe_doc(vm, a):
            p = a[0].p
            if p["i"] < 0:
                return -1
            return p["list"][p["i"]][0] if p["i"] < len(p["list"]) else NO_MORE_DOCS

        def pe_adv(vm, a):
            p = a[0].p
            while True:
                p["i"] += 1
                if p["i"] >= len(p["list"]):
                    return NO_MORE_DOCS
                if p["list"][p["i"]][0] >= a[1]:
                    p["pi"] = 0
                    return p["list"][p["i"]][0]

        def pe_freq(vm, a):
            p = a[0].p
            return len(p["list"][p["i"]][1])

        def pe_nextpos(vm, a):
            p = a[0].p
            pos = p["list"][p["i"]][1][p["pi"]]
            p["pi"] += 1
            return pos

        vm.define_class("stub/Postings", P + "index/PostingsEnum", methods={
            ("nextDoc", "()I"): pe_next, ("docID", "()I"): pe_doc, ("advance", "(I)I"): pe_adv,
            ("cost", "()J"): lambda vm, a: len(a[0].p["list"]),
            ("freq", "()I"): pe_freq, ("nextPosition", "()I"): pe_nextpos,
            ("startOffset", "()I"): lambda vm, a: -1, ("endOffset", "()I"): lambda vm, a: -1,
            ("getPayload", "()L%sutil/BytesRef;" % P): lambda vm, a: None,
        })

        def te_seek(vm, a):
            t = bref_bytes(a[1])
            a[0].p = t if t in me.postings else None
            return 1 if a[0].p is not None else 0

        def te_postings(vm, a):
            o = vm.new("stub/Postings")
            vm.call_special(vm.load(P + "index/PostingsEnum"), "<init>", "()V", o)
            o.p = {"list": me.postings[a[0].p], "i": -1, "pi": 0}
            return o

        def te_term(vm, a):
            return vm.new_init(P + "util/BytesRef", "([B)V", jbytes(a[0].p))

        vm.define_class("stub/TermsEnum", P + "index/TermsEnum", methods={
            ("seekExact", "(L%sutil/BytesRef;)Z" % P): te_seek,
            ("term", "()L%sutil/BytesRef;" % P): te_term,
            ("docFreq", "()I"): lambda vm, a: len(me.postings[a[0].p]),
            ("totalTermFreq", "()J"): lambda vm, a: sum(len(x[1]) for x in me.postings[a[0].p]),
            ("postings", "(L%sindex/PostingsEnum;I)L%sindex/PostingsEnum;" % (P, P)): te_postings,
        })

        def terms_iter(vm, a):
            o = vm.new("stub/TermsEnum")
            vm.call_special(vm.load(P + "index/TermsEnum"), "<init>", "()V", o)
            return o

        vm.define_class("stub/Terms", P + "index/Terms", methods={
            ("iterator", "()L%sindex/TermsEnum;" % P): terms_iter,
            ("size", "()J"): lambda vm, a: len(me.postings),
            ("getSumTotalTermFreq", "()J"): lambda vm, a: sum(me.doc_len),
            ("getSumDocFreq", "()J"): lambda vm, a: sum(len(v) for v in me.postings.values()),
            ("getDocCount", "()I"): lambda vm, a: sum(1 for n in me.doc_len if n > 0),
            ("hasFreqs", "()Z"): lambda vm, a: 1, ("hasOffsets", "()Z"): lambda vm, a: 0,
            ("hasPositions", "()Z"): lambda vm, a: 1, ("hasPayloads", "()Z"): lambda vm, a: 0,
        })

        def fields_terms(vm, a):
            return me.terms_obj if a[1] == me.field else None

        vm.define_class("stub/Fields", P + "index/Fields", methods={
            ("terms", "(Ljava/lang/String;)L%sindex/Terms;" % P): fields_terms,
            ("size", "()I"): lambda vm, a: 1,
        })
        vm.define_class("stub/Norms", P + "index/NumericDocValues", methods={
            ("get", "(I)J"): lambda vm, a: me.norm_long[a[1]],
        })
        vm.define_class("stub/Reader", P + "index/LeafReader", methods={
            ("fields", "()L%sindex/Fields;" % P): lambda vm, a: me.fields_obj,
            ("getNormValues", "(Ljava/lang/String;)L%sindex/NumericDocValues;" % P): lambda vm, a: me.norms_obj if a[1] == me.field else None,
            ("getLiveDocs", "()L%sutil/Bits;" % P): lambda vm, a: None,
            ("numDocs", "()I"): lambda vm, a: len(me.doc_len),
            ("maxDoc", "()I"): lambda vm, a: len(me.doc_len),
        })

    # ------------------------------------------------------------------ index
    def index(self, docs):
        """docs: list of documents, each a list of term bytes in token order (position = ordinal)."""
        vm = self.vm
        self.postings = {}
        self.doc_len = []
        for d, toks in enumerate(docs):
            self.doc_len.append(len(toks))
            for pos, t in enumerate(toks):
                pl = self.postings.setdefault(bytes(t), [])
                if not pl or pl[-1][0] != d:
                    pl.append((d, []))
                pl[-1][1].append(pos)
        # norms: Similarity#computeNorm(FieldInvertState) from the jar (TFIDFSimilarity#computeNorm /
        # BM25Similarity#computeNorm); Lucene's norms consumer stores the long, its low byte is the norm byte
        self.norm_long = []
        for n in self.doc_len:
            fis = vm.new_init(P + "index/FieldInvertState", "(Ljava/lang/String;)V", self.field)
            vm.call_virtual(fis, "setLength", "(I)V", n)
            vm.call_virtual(fis, "setBoost", "(F)V", 1.0)
            self.norm_long.append(vm.call_virtual(self.similarity, "computeNorm", "(L%sindex/FieldInvertState;)J" % P, fis))
        self.terms_obj = vm.new("stub/Terms")
        vm.call_special(vm.load(P + "index/Terms"), "<init>", "()V", self.terms_obj)
        self.fields_obj = vm.new("stub/Fields")
        vm.call_special(vm.load(P + "index/Fields"), "<init>", "()V", self.fields_obj)
        self.norms_obj = vm.new("stub/Norms")
        vm.call_special(vm.load(P + "index/NumericDocValues"), "<init>", "()V", self.norms_obj)
        self.reader = vm.new("stub/Reader")
        vm.call_special(vm.load(P + "index/LeafReader"), "<init>", "()V", self.reader)
        # Classifier.java:103-106: new IndexSearcher(reader); setSimilarity(similarity)
        self.searcher = vm.new_init(P + "search/IndexSearcher", "(L%sindex/IndexReader;)V" % P, self.reader)
        vm.call_virtual(self.searcher, "setSimilarity", "(L%ssearch/similarities/Similarity;)V" % P, self.similarity)
        return self

    def norm_bytes(self):
        return [n & 0xFF for n in self.norm_long]

    # ------------------------------------------------------------------ query (Classifier.java:113-261)
    def term(self, t: bytes):
        vm = self.vm
        br = vm.new_init(P + "util/BytesRef", "([B)V", jbytes(t))
        br = vm.call_static(P + "util/BytesRef", "deepCopyOf", "(L%sutil/BytesRef;)L%sutil/BytesRef;" % (P, P), br)
        return vm.new_init(P + "index/Term", "(Ljava/lang/String;L%sutil/BytesRef;)V" % P, self.field, br)

    def build_query(self, clauses, msm_fraction: float):
        """clauses: ('term', t) | ('phrase', t0, t1, slop) in clause order → the BooleanQuery of createQuery()."""
        vm = self.vm
        BQB = P + "search/BooleanQuery$Builder"
        should = vm.load(P + "search/BooleanClause$Occur")
        vm.init_class(should)
        SHOULD = should.statics["SHOULD"]
        q = vm.new_init(BQB, "()V")
        vm.call_virtual(q, "setDisableCoord", "(Z)L%s;" % BQB, 0)
        for c in clauses:
            if c[0] == "term":
                sub = vm.new_init(P + "search/TermQuery", "(L%sindex/Term;)V" % P, self.term(c[1]))
            else:
                PQB = P + "search/PhraseQuery$Builder"
                pq = vm.new_init(PQB, "()V")
                vm.call_virtual(pq, "setSlop", "(I)L%s;" % PQB, c[3])
                vm.call_virtual(pq, "add", "(L%sindex/Term;)L%s;" % (P, PQB), self.term(c[1]))
                vm.call_virtual(pq, "add", "(L%sindex/Term;)L%s;" % (P, PQB), self.term(c[2]))
                sub = vm.call_virtual(pq, "build", "()L%ssearch/PhraseQuery;" % P)
            vm.call_virtual(q, "add", "(L%ssearch/Query;L%ssearch/BooleanClause$Occur;)L%s;" % (P, P, BQB), sub, SHOULD)
        qc = vm.call_virtual(q, "build", "()L%ssearch/BooleanQuery;" % P)
        # createQuery(): second builder, msm = (int)(minShouldMatch * clauses().size())
        b2 = vm.new_init(BQB, "()V")
        vm.call_virtual(b2, "setDisableCoord", "(Z)L%s;" % BQB, vm.call_virtual(qc, "isCoordDisabled", "()Z"))
        ncl = vm.call_virtual(vm.call_virtual(qc, "clauses", "()Ljava/util/List;"), "size", "()I")
        msm = J._f2i(msm_fraction * float(ncl), 31)  # Java: (int)(double * int)
        vm.call_virtual(b2, "setMinimumNumberShouldMatch", "(I)L%s;" % BQB, msm)
        it = vm.call_virtual(qc, "iterator", "()Ljava/util/Iterator;")
        while vm.call_virtual(it, "hasNext", "()Z"):
            cl = vm.call_virtual(it, "next", "()Ljava/lang/Object;")
            vm.call_virtual(b2, "add", "(L%ssearch/BooleanClause;)L%s;" % (P, BQB), cl)
        return vm.call_virtual(b2, "build", "()L%ssearch/BooleanQuery;" % P), msm

    # ------------------------------------------------------------------ search (Classifier.java:350-366)
    def search(self, query, n: int = 10):
        vm = self.vm
        coll = vm.call_static(P + "search/TopScoreDocCollector", "create", "(I)L%ssearch/TopScoreDocCollector;" % P, n)
        vm.call_virtual(self.searcher, "search", "(L%ssearch/Query;L%ssearch/Collector;)V" % (P, P), query, coll)
        td = vm.call_virtual(coll, "topDocs", "()L%ssearch/TopDocs;" % P)
        hits = []
        for sd in td.f[P + "search/TopDocs.scoreDocs"].a:
            hits.append((sd.f[P + "search/ScoreDoc.doc"], sd.f[P + "search/ScoreDoc.score"]))
        max_score = vm.call_virtual(td, "getMaxScore", "()F")
        total = td.f[P + "search/TopDocs.totalHits"]
        return {"hits": hits, "max_score": max_score, "total_hits": total}

    def weights(self, query):
        """The normalized per-clause weight values (IDFStats.value / BM25Stats.weight), in clause order."""
        vm = self.vm
        w = vm.call_virtual(self.searcher, "createNormalizedWeight", "(L%ssearch/Query;Z)L%ssearch/Weight;" % (P, P), query, 1)
        out = []
        if not vm.instance_of(w, P + "search/BooleanWeight"):
            subs = [w]
        else:
            subs = list(w.f[P + "search/BooleanWeight.weights"].p)
        for sw in subs:
            if vm.instance_of(sw, P + "search/TermQuery$TermWeight"):
                stats = sw.f[P + "search/TermQuery$TermWeight.stats"]
            else:
                stats = sw.f[P + "search/PhraseQuery$PhraseWeight.stats"]
            if self.scoring == "bm25":
                out.append(stats.f[P + "search/similarities/BM25Similarity$BM25Stats.weight"])
            else:
                out.append(stats.f[P + "search/similarities/TFIDFSimilarity$IDFStats.value"])
        return out

    # ------------------------------------------------------------------ SloppyPhraseScorer#phraseFreq on position lists
    def phrase_freq(self, p0, p1, slop: int, same_term: bool = False):
        """One document holding t0 at positions p0 and t1 at p1 (same_term: a single term at positions p0; p1 is
        ignored) → `SloppyPhraseScorer#sloppyFreq()` after `nextDoc()`, or None when the phrase does not match."""
        vm = self.vm
        self.field = "sequence"
        n = max(list(p0) + ([] if same_term else list(p1))) + 1
        toks = [b"_"] * n
        for p in p0:
            toks[p] = b"a"
        if not same_term:
            for p in p1:
                assert toks[p] == b"_", "position used twice"
                toks[p] = b"b"
        self.index([toks])
        PQB = P + "search/PhraseQuery$Builder"
        pq = vm.new_init(PQB, "()V")
        vm.call_virtual(pq, "setSlop", "(I)L%s;" % PQB, slop)
        vm.call_virtual(pq, "add", "(L%sindex/Term;)L%s;" % (P, PQB), self.term(b"a"))
        vm.call_virtual(pq, "add", "(L%sindex/Term;)L%s;" % (P, PQB), self.term(b"a" if same_term else b"b"))
        q = vm.call_virtual(pq, "build", "()L%ssearch/PhraseQuery;" % P)
        w = vm.call_virtual(self.searcher, "createNormalizedWeight", "(L%ssearch/Query;Z)L%ssearch/Weight;" % (P, P), q, 1)
        leaves = vm.call_virtual(vm.call_virtual(self.reader, "getContext", "()L%sindex/LeafReaderContext;" % P),
                                 "leaves", "()Ljava/util/List;")
        ctx = vm.call_virtual(leaves, "get", "(I)Ljava/lang/Object;", 0)
        sc = vm.call_virtual(w, "scorer", "(L%sindex/LeafReaderContext;)L%ssearch/Scorer;" % (P, P), ctx)
        if sc is None:
            return None
        assert sc.cls.name == P + "search/SloppyPhraseScorer", sc.cls.name
        d = vm.call_virtual(sc, "nextDoc", "()I")
        if d == NO_MORE_DOCS:
            return None
        return vm.call_virtual(sc, "sloppyFreq", "()F")
