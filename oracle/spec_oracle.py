"""Spec oracle -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A slow, literal, pure-Python restatement of the BioSpectra index + classify hot
path and of the Lucene 5.3.1 arithmetic it delegates to.  Written to be
"obviously the same as the reference", not to be fast: use it on small cases
only.  The fast C oracle (oracle/bsp_oracle.c) and the CUDA path are both
checked against it.

PARITY PINNED BY EXECUTION: the reference ships no tests / golden outputs and no
JVM exists here, so oracle/minijvm.py executes the reference's own class files
(libs/lucene-core-5.3.1.jar, fasta-utils-1.1.jar) and tests/golden/lucene_*.json
holds what they returned (SmallFloat, DefaultSimilarity / BM25Similarity,
SloppyPhraseScorer#phraseFreq, whole IndexSearcher.search runs, FASTAReader);
tests/test_lucene_golden.py requires this oracle to reproduce every vector bit for
bit.  The two Java *source* files on the path (KmerSequenceTokenizer.java,
Classifier.java) are restated below line by line (no compiler here); SURVEY.md
Appendix B's tokenizer / clause-count rows check them (tests/test_oracle_kat.py).

Every function cites the reference file:line (relative to /root/reference) or
the Lucene class#method it follows.  All float32 arithmetic goes through
numpy.float32 scalars so each operation rounds exactly like a Java `float`.
"""
from __future__ import annotations

import base64
import json
import math
import re
import struct
from fractions import Fraction

import numpy as np

f32 = np.float32

NAIVE_KMER, CHAIN_PROXIMITY, PAIRED_PROXIMITY = 0, 1, 2
ALGORITHMS = {"NAIVE_KMER": 0, "CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}
TFIDF, BM25 = 0, 1


# ----------------------------------------------------------------------------
# A.1  FASTA reading: fasta-utils-1.1.jar!org/yeastrc/fasta/FASTAReader#readNext
# ----------------------------------------------------------------------------
def split_lines(text: str):
    """java.io.BufferedReader#readLine: \\n, \\r\\n and lone \\r all end a line."""
    return text.replace("\r\n", "\n").replace("\r", "\n").split("\n")


_HDR_NAME = re.compile(r"(\S+)\s*", re.ASCII)               # FASTAHeader.<init>: ^(\S+)\s*$
_HDR_NAME_DESC = re.compile(r"(\S+)\s+(\S+.*)", re.ASCII)    #                     ^(\S+)\s+(\S+.*)$


def _check_header(line: str):
    """FASTAReader#readNext: the header line minus '>' is split on Ctrl-A and every part must parse as a
    FASTAHeader (name [description]); e.g. '> x' or a bare '>' raise FASTADataErrorException."""
    body = line[1:]
    if "\x01" not in body:
        parts = [body]                    # String#split: no match -> the input itself (even when empty)
    else:
        parts = body.split("\x01")
        while parts and parts[-1] == "":
            parts.pop()                   # trailing empty strings are removed
    for p in parts:
        if not (_HDR_NAME.fullmatch(p) or _HDR_NAME_DESC.fullmatch(p)):
            raise ValueError("FASTADataErrorException: Could not parse FASTA header line: \"%s\"" % p)


def iter_fasta(text: str):
    """fasta-utils-1.1.jar!org/yeastrc/fasta/FASTAReader#readNext, statement for statement (pinned by
    tests/golden/fasta_reader.json, produced by executing that class file).  Yields (headerLine incl. '>', sequence)
    and raises where the reference throws, after the entries it had already returned."""
    lines = split_lines(text)
    if lines and lines[-1] == "":
        lines.pop()                       # no phantom line after the last newline
    it = iter(lines)
    last = next(it, None)                 # lineNumber == 0: read the first line
    while True:
        line = last
        if line is None:
            return
        if not line.startswith(">"):
            raise ValueError("FASTADataErrorException: Expected header line, but line did not start with \">\".")
        header = line
        _check_header(line)
        line = next(it, None)
        while True:                       # while (line.startsWith(";")) — dereferences null at EOF
            if line is None:
                raise ValueError("NullPointerException: header line at end of file")
            if not line.startswith(";"):
                break
            line = next(it, None)
        if line.startswith(">"):
            raise ValueError("FASTADataErrorException: Did not get a sequence line after a header line")
        seq = []
        while line is not None and not line.startswith(">"):
            if not line.startswith(";"):
                seq.append(line.upper())  # no per-line trim: interior blanks survive
            line = next(it, None)
        last = line
        yield (header, java_trim("".join(seq)))


def read_fasta(text: str):
    return list(iter_fasta(text))


def java_trim(s: str) -> str:
    """String#trim: strips chars <= U+0020 from both ends."""
    a, b = 0, len(s)
    while a < b and s[a] <= " ":
        a += 1
    while b > a and s[b - 1] <= " ":
        b -= 1
    return s[a:b]


# ----------------------------------------------------------------------------
# A.2  tokenizer + term encoding
# ----------------------------------------------------------------------------
def tokenize(seq: str, k: int, skips: int):
    """lucene/KmerSequenceTokenizer.java:115-154 (buffer refills are transparent
    for reads <= 4096 chars; the index analyzer always uses skips=0,
    lucene/KmerIndexAnalyzer.java:48)."""
    if k < 1 or skips < 0:
        raise ValueError("bad k/skips")
    toks = []
    pos = 0
    n = len(seq)
    while True:                                   # one incrementToken() call
        cur_skip = skips
        emitted = False
        while True:
            if pos + k > n:
                break
            w = seq[pos:pos + k]
            if all(c in "ATGC" for c in w):
                toks.append((w, pos))
                pos += 1 + cur_skip
                emitted = True
                break
            cur_skip = 0
            pos += 1 + cur_skip
        if not emitted:
            return toks


_COMP = {}
for _c in "ABCDEFGHIJKLMNOPQRSTUVWXYZ":
    _COMP[_c] = " "
_COMP.update({"A": "T", "T": "A", "C": "G", "G": "C", "N": "N", "X": "X"})


def reverse_complement(seq: str) -> str:
    """utils/SequenceHelper.java:26-35,85-91.  Any char outside 'A'..'Z' indexes
    outside the 26-entry LUT -> ArrayIndexOutOfBoundsException (IndexError here)."""
    out = []
    for ch in reversed(seq):
        if ch not in _COMP:
            raise IndexError("ArrayIndexOutOfBounds in ComplementCharLUT: %r" % ch)
        out.append(_COMP[ch])
    return "".join(out)


_BITS = {"A": 0, "C": 1, "G": 2, "T": 3}


def pack_kmer(kmer: str) -> int:
    """Integer view of the 2-bit code (first base most significant)."""
    v = 0
    for c in kmer:
        v = (v << 2) | _BITS[c]
    return v


def compress(kmer: str) -> bytes:
    """utils/SequenceHelper.java:225-271: 4 bases/byte, MSB first, zero tail."""
    out = bytearray((len(kmer) + 3) // 4)
    for i, c in enumerate(kmer):
        out[i // 4] |= _BITS[c] << (6 - 2 * (i % 4))
    return bytes(out)


def canonical(kmer: str, min_strand: bool) -> str:
    """lucene/SequenceCompressFilter.java:56-68 (String#compareTo, A<C<G<T)."""
    if not min_strand:
        return kmer
    rc = reverse_complement(kmer)
    return rc if kmer > rc else kmer


def term_text(kmer: str, min_strand: bool = False) -> str:
    """lucene/SequenceCompressFilter.java:70-86: Base64(compress(kmer))."""
    return base64.b64encode(compress(canonical(kmer, min_strand))).decode("ascii")


# ----------------------------------------------------------------------------
# SmallFloat + similarities (A.5, A.6)
# ----------------------------------------------------------------------------
def float_to_raw_bits(x) -> int:
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


def bits_to_float(b: int):
    return f32(struct.unpack("<f", struct.pack("<I", b & 0xFFFFFFFF))[0])


def float_to_byte315(x) -> int:
    """lucene-core!util/SmallFloat#floatToByte315."""
    bits = float_to_raw_bits(x)
    s = bits >> 21
    if s <= ((63 - 15) << 3):
        return 0 if bits <= 0 else 1
    if s >= ((63 - 15) << 3) + 0x100:
        return 0xFF
    return (s - ((63 - 15) << 3)) & 0xFF


def byte315_to_float(b: int):
    """lucene-core!util/SmallFloat#byte315ToFloat."""
    if b == 0:
        return f32(0.0)
    return bits_to_float(((b & 0xFF) << 21) + ((63 - 15) << 24))


NORM_TABLE = [byte315_to_float(b) for b in range(256)]


def norm_byte(num_tokens: int) -> int:
    """TFIDFSimilarity#computeNorm -> DefaultSimilarity#lengthNorm/#encodeNormValue
    (BM25Similarity#computeNorm/#encodeNormValue store the same byte)."""
    if num_tokens == 0:
        # lengthNorm: (float)(1.0 / Math.sqrt(0)) = +Inf -> 0xFF
        return float_to_byte315(f32(np.inf))
    return float_to_byte315(f32(1.0) * f32(1.0 / math.sqrt(num_tokens)))


def idf_tfidf(df: int, max_doc: int):
    """DefaultSimilarity#idf."""
    return f32(math.log(float(max_doc) / float(df + 1)) + 1.0)


def idf_bm25(df: int, max_doc: int):
    """BM25Similarity#idf."""
    return f32(math.log(1.0 + (float(max_doc - df) + 0.5) / (float(df) + 0.5)))


def tf_float(freq):
    """DefaultSimilarity#tf: (float)Math.sqrt(freq)."""
    return f32(math.sqrt(float(freq)))


def sloppy_weight(distance: int):
    """DefaultSimilarity#sloppyFreq / BM25Similarity#sloppyFreq: 1.0f/(distance+1)."""
    return f32(1.0) / f32(distance + 1)


# ----------------------------------------------------------------------------
# A.3  document model + inversion
# ----------------------------------------------------------------------------
class Doc:
    __slots__ = ("filename", "header", "direction", "taxonomy", "ntok", "norm")

    def __init__(self, filename, header, direction, taxonomy):
        self.filename, self.header, self.direction, self.taxonomy = filename, header, direction, taxonomy
        self.ntok = 0
        self.norm = 0


class Index:
    """term (packed canonical k-mer int) -> {doc: [positions]}; positions are token
    ordinals (FieldInvertState position, posIncr always 1)."""

    def __init__(self, k: int, min_strand: bool):
        self.k, self.min_strand = k, min_strand
        self.docs: list[Doc] = []
        self.postings: dict[int, dict[int, list[int]]] = {}

    @property
    def max_doc(self):
        return len(self.docs)

    def _add_doc(self, doc: Doc, sequence: str):
        did = len(self.docs)
        self.docs.append(doc)
        toks = tokenize(sequence, self.k, 0)          # KmerIndexAnalyzer: skips 0
        for pos, (kmer, _off) in enumerate(toks):
            t = pack_kmer(canonical(kmer, self.min_strand))
            self.postings.setdefault(t, {}).setdefault(did, []).append(pos)
        doc.ntok = len(toks)
        doc.norm = norm_byte(len(toks))

    def add_entry(self, filename: str, header: str, taxonomy: str, sequence: str):
        """index/Indexer.java:178-232.  header has its leading '>' removed by the
        caller (:167-170).  Canonical doc ids: dense, in add order (forward before reverse)."""
        if self.min_strand:
            self._add_doc(Doc(filename, header, "min_strand", taxonomy), sequence)
            return
        self._add_doc(Doc(filename, header, "forward", taxonomy), sequence)
        try:
            rc = reverse_complement(sequence)
        except IndexError:
            # exception swallowed at Indexer.java:227-229: the forward doc stays, the
            # reverse doc is never added (so it does not count in maxDoc either).
            return
        self._add_doc(Doc(filename, header, "reverse", taxonomy), rc)

    def df(self, t: int) -> int:
        return len(self.postings.get(t, ()))

    def sum_total_term_freq(self) -> int:
        return sum(d.ntok for d in self.docs)


# ----------------------------------------------------------------------------
# A.4  query construction (classify/Classifier.java:113-261)
# ----------------------------------------------------------------------------
def build_clauses(tokens, algorithm: int, min_strand: bool):
    """tokens: [(kmer, offset)].  Returns list of ('T', term) / ('P', t0, t1, slop)
    or None when the read has < 2 tokens (Classifier.java:223-227 -> NPE -> dropped)."""
    n = len(tokens)
    if n <= 1:
        return None
    terms = [pack_kmer(canonical(km, min_strand)) for km, _ in tokens]
    offs = [o for _, o in tokens]
    clauses = []
    if algorithm == NAIVE_KMER:                       # :113-118
        clauses = [("T", t) for t in terms]
    elif algorithm == CHAIN_PROXIMITY:                # :120-158
        for i in range(1, n):
            d = offs[i] - offs[i - 1]
            if d > 0:
                clauses.append(("P", terms[i - 1], terms[i], d + 1))
    elif algorithm == PAIRED_PROXIMITY:               # :160-200
        for j in range(0, n - 1, 2):
            d = offs[j + 1] - offs[j]
            if d > 0:
                clauses.append(("P", terms[j], terms[j + 1], d + 1))
        if n % 2 == 1:
            clauses.append(("T", terms[n - 1]))
    else:
        raise ValueError("algorithm")
    return clauses


# ----------------------------------------------------------------------------
# A.7  SloppyPhraseScorer, literal (2 phrase positions, offsets 0 and 1)
# ----------------------------------------------------------------------------
class _PP:
    """lucene-core!search/PhrasePositions."""

    def __init__(self, plist, offset, ordinal):
        self.plist, self.offset, self.ord = plist, offset, ordinal
        self.position = 0
        self.count = 0
        self.idx = 0
        self.rpt_group = -1
        self.rpt_ind = 0

    def first_position(self):
        self.count = len(self.plist)
        self.idx = 0
        self.next_position()

    def next_position(self):
        if self.count > 0:
            self.count -= 1
            self.position = self.plist[self.idx] - self.offset
            self.idx += 1
            return True
        self.count -= 1
        return False

    def key(self):                                    # PhraseQueue#lessThan
        return (self.position, self.offset, self.ord)


class _SloppyScorer:
    def __init__(self, p0, p1, slop, same_term):
        self.pps = [_PP(p0, 0, 0), _PP(p1, 1, 1)]
        self.slop = slop
        self.has_rpts = same_term
        self.pq = []
        self.end = 0
        self.rg = None

    # --- queue helpers (a 2-element binary heap == pick-the-min list)
    def _pop(self):
        m = min(self.pq, key=_PP.key)
        self.pq.remove(m)
        return m

    def _top(self):
        return min(self.pq, key=_PP.key)

    def advance_pp(self, pp):                         # #advancePP
        if not pp.next_position():
            return False
        if pp.position > self.end:
            self.end = pp.position
        return True

    @staticmethod
    def tp_pos(pp):                                   # #tpPos
        return pp.position + pp.offset

    def collide(self, pp):                            # #collide
        tp = self.tp_pos(pp)
        for other in self.rg:
            if other is not pp and self.tp_pos(other) == tp:
                return other.rpt_ind
        return -1

    @staticmethod
    def lesser(a, b):                                 # #lesser
        if a.position < b.position or (a.position == b.position and a.offset < b.offset):
            return a
        return b

    def advance_rpts(self, pp):                       # #advanceRpts
        if pp.rpt_group < 0:
            return True
        bits = set()
        k0 = pp.rpt_ind
        while True:
            k = self.collide(pp)
            if k < 0:
                break
            pp = self.lesser(pp, self.rg[k])          # rebinds the LOCAL only
            if not self.advance_pp(pp):
                return False
            if k != k0:
                bits.add(k)
        # pps advanced while inside pq are popped and re-added (heap repair): no-op
        # for a min-by-key list.
        return True

    def init_positions(self):                         # #initPhrasePositions
        self.end = -(1 << 31)
        for pp in self.pps:                           # placeFirstPositions
            pp.first_position()
        if self.has_rpts:
            # gatherRptGroups (single-term repeats): both pps share the term and have
            # the same tpPos at their first position -> one group, sorted by offset.
            self.rg = sorted(self.pps, key=lambda p: p.offset)
            for i, pp in enumerate(self.rg):
                pp.rpt_group = 0
                pp.rpt_ind = i
            # advanceRepeatGroups (!hasMultiTermRpts): pp[i] advanced i times
            for i in range(1, len(self.rg)):
                for _ in range(i):
                    if not self.rg[i].next_position():
                        return False
        self.pq = []                                  # fillQueue / initSimple
        for pp in self.pps:
            if pp.position > self.end:
                self.end = pp.position
            self.pq.append(pp)
        return True

    def phrase_freq(self):                            # #phraseFreq
        if not self.init_positions():
            return f32(0.0)
        freq = f32(0.0)
        pp = self._pop()
        match_length = self.end - pp.position
        nxt = self._top().position
        while self.advance_pp(pp):
            if self.has_rpts and not self.advance_rpts(pp):
                break
            if pp.position > nxt:
                if match_length <= self.slop:
                    freq = f32(freq + sloppy_weight(match_length))
                self.pq.append(pp)
                pp = self._pop()
                nxt = self._top().position
                match_length = self.end - pp.position
            else:
                ml2 = self.end - pp.position
                if ml2 < match_length:
                    match_length = ml2
        if match_length <= self.slop:
            freq = f32(freq + sloppy_weight(match_length))
        return freq


def sloppy_freq(p0, p1, slop: int, same_term: bool = False):
    """Phrase frequency of (t0@0, t1@1) with the given slop in one document whose
    position lists are p0, p1 (both non-empty).  same_term: t0 == t1."""
    return _SloppyScorer(list(p0), list(p1), slop, same_term).phrase_freq()


# ----------------------------------------------------------------------------
# A.5 / A.6 / A.8  weights, per-document score, collector
# ----------------------------------------------------------------------------
def clause_idf(index: Index, clause, scoring: int):
    max_doc = index.max_doc
    idf = idf_tfidf if scoring == TFIDF else idf_bm25
    if clause[0] == "T":
        return idf(index.df(clause[1]), max_doc)
    # TFIDFSimilarity#idfExplain(stats, TermStatistics[]): float sum, t0 first
    acc = f32(0.0)
    for t in (clause[1], clause[2]):
        acc = f32(acc + idf(index.df(t), max_doc))
    return acc


def clause_freqs(index: Index, clause):
    """{doc: freq(float32)} for docs where the clause matches."""
    out = {}
    if clause[0] == "T":
        for d, pl in index.postings.get(clause[1], {}).items():
            out[d] = f32(len(pl))                     # TermScorer: freq as float
        return out
    _, t0, t1, slop = clause
    a = index.postings.get(t0, {})
    b = index.postings.get(t1, {})
    for d in a.keys() & b.keys():                     # ConjunctionDISI
        fr = sloppy_freq(a[d], b[d], slop, t0 == t1)
        if fr != 0:
            out[d] = fr
    return out


def query_norm(sum_sq):
    """DefaultSimilarity#queryNorm: (float)(1.0 / Math.sqrt(sumOfSquaredWeights)); IndexSearcher#createNormalizedWeight
    replaces Inf/NaN by 1.0f."""
    with np.errstate(divide="ignore", invalid="ignore"):
        qn = f32(1.0 / math.sqrt(float(sum_sq))) if sum_sq > 0 else f32(np.inf)
    if not np.isfinite(qn):
        qn = f32(1.0)
    return qn


def raw_query_norm(sum_sq):
    """DefaultSimilarity#queryNorm alone (no Inf/NaN replacement)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return f32(1.0 / math.sqrt(float(sum_sq))) if sum_sq > 0 else (f32(np.inf) if sum_sq == 0 else f32(np.nan))


def tfidf_sum_sq(idfs):
    """BooleanWeight#getValueForNormalization over IDFStats#getValueForNormalization (queryWeight = idf * 1.0f)."""
    sum_sq = f32(0.0)
    for w in idfs:
        qw = f32(w * f32(1.0))
        sum_sq = f32(sum_sq + f32(qw * qw))
    return f32(sum_sq * f32(f32(1.0) * f32(1.0)))


def tfidf_values(idfs):
    """TFIDFSimilarity$IDFStats#normalize: value = (queryWeight * (queryNorm * topLevelBoost)) * idf."""
    qn = query_norm(tfidf_sum_sq(idfs))
    return [f32(f32(f32(w * f32(1.0)) * f32(qn * f32(1.0))) * w) for w in idfs]


def tfidf_coord(i: int, n_clauses: int):
    """BooleanWeight#coord / DefaultSimilarity#coord: 0 overlap -> 0; one clause -> 1; else (float)i / (float)n."""
    if i == 0:
        return f32(0.0)
    if n_clauses == 1:
        return f32(1.0)
    return f32(i) / f32(n_clauses)


def avg_field_length(sum_total_term_freq: int, max_doc: int):
    """BM25Similarity#avgFieldLength."""
    return f32(1.0) if sum_total_term_freq <= 0 else f32(float(sum_total_term_freq) / float(max_doc))


def bm25_cache(avgdl):
    """BM25Similarity#computeWeight: cache[x] = k1 * ((1 - b) + b * decodeNormValue(x) / avgdl),
    decodeNormValue(x) = NORM_TABLE[x] = 1 / (byte315ToFloat(x)^2)."""
    k1, b = f32(1.2), f32(0.75)
    cache = []
    for x in range(256):
        nv = NORM_TABLE[x]
        with np.errstate(divide="ignore"):
            ln = f32(1.0) / f32(nv * nv)
        cache.append(f32(k1 * f32(f32(f32(1.0) - b) + f32(f32(b * ln) / avgdl))))
    return cache


def search(index: Index, sequence: str, *, skips: int, algorithm: int, scoring: int,
           msm_ratio: float, top_n: int = 10, check_exact: bool = True, details: dict | None = None):
    """classify/Classifier.java:341-366 + IndexSearcher/BooleanWeight/BooleanScorer/
    TopScoreDocCollector.  Returns None if the read is dropped (< 2 tokens or more
    than 10000 clauses), else (hits, n_clauses, msm) with hits = [(doc, score f32)]
    sorted score desc, doc asc, at most top_n."""
    if not sequence:
        raise ValueError("sequence is null or empty")
    tokens = tokenize(sequence, index.k, skips)
    clauses = build_clauses(tokens, algorithm, index.min_strand)
    if clauses is None:
        return None
    n_clauses = len(clauses)
    if n_clauses > 10000:                             # BooleanQuery.setMaxClauseCount(10000)
        return None
    msm = int(msm_ratio * n_clauses)                  # Classifier.java:257
    idfs = [clause_idf(index, c, scoring) for c in clauses]
    if scoring == TFIDF:
        values = tfidf_values(idfs)
        coord = [tfidf_coord(i, n_clauses) for i in range(n_clauses + 1)]
    else:
        values = [f32(f32(w * f32(1.0)) * f32(1.0)) for w in idfs]     # BM25Stats#normalize: idf * queryBoost * topLevelBoost
        coord = [f32(0.0)] + [f32(1.0)] * n_clauses
        cache = bm25_cache(avg_field_length(index.sum_total_term_freq(), index.max_doc))
    if details is not None:
        details.update(values=values, n_clauses=n_clauses, msm=msm)

    acc: dict[int, list] = {}                         # doc -> [exact sum, double sum, count]
    for c, val in zip(clauses, values):
        for d, fr in clause_freqs(index, c).items():
            nb = index.docs[d].norm
            if scoring == TFIDF:
                s = f32(f32(tf_float(fr) * val) * NORM_TABLE[nb])
            else:
                wv = f32(val * f32(f32(1.2) + f32(1.0)))
                s = f32(f32(wv * fr) / f32(fr + cache[nb]))
            e = acc.setdefault(d, [Fraction(0), 0.0, 0])
            e[0] += Fraction(float(s))
            e[1] += float(s)
            e[2] += 1
    need = max(1, msm)
    hits = []
    for d, (exact, dsum, m) in acc.items():
        if m < need:
            continue
        if check_exact:
            # double accumulation must be exact for the result to be order-independent
            assert Fraction(dsum) == exact, "double accumulation not exact"
        hits.append((d, f32(f32(dsum) * coord[m])))
    hits.sort(key=lambda h: (-float(h[1]), h[0]))     # TopScoreDocCollector order
    if details is not None:
        details["total_hits"] = len(hits)
    return hits[:top_n], n_clauses, msm


# ----------------------------------------------------------------------------
# A.8 / A.9  result set, label, JSON
# ----------------------------------------------------------------------------
def _ranked(lineage_json: str):
    """beans/TaxonTreeDescription.java:33-98: list of taxa whose rank is not
    null/""/"no rank" (case-insensitive), in file order (leaf -> root)."""
    js = lineage_json
    if js.startswith("[") and js.endswith("]"):
        tree = json.loads(js)
    elif js.startswith("{") and js.endswith("}"):
        tree = json.loads(js).get("tree", [])
    else:
        raise IOError("cannot create instance from wrong json format")
    out = []
    for t in tree:
        r = t.get("rank")
        if r is not None and r != "" and r.lower() != "no rank":
            out.append(t)
    return out


def label(entries):
    """classify/Classifier.java:263-339.  entries: list of dicts with key
    'taxon_hierarchy'.  Returns (type, taxon_rank, taxon_name)."""
    if not entries:
        return "UNKNOWN", "unknown", ""
    if len(entries) == 1:
        th = entries[0]["taxon_hierarchy"]
        if th:
            r = _ranked(th)
            if r:
                return "CLASSIFIED", r[0].get("rank"), r[0].get("name")
        return "CLASSIFIED", "unknown", ""
    trees = []
    for e in entries:
        th = e["taxon_hierarchy"]
        if not th:
            return "VAGUE", "unknown", ""
        trees.append(_ranked(th))
    for tax in trees[0]:
        if all(any(t2.get("taxid", 0) == tax.get("taxid", 0) for t2 in other) for other in trees[1:]):
            return "CLASSIFIED", tax.get("rank"), tax.get("name")
    return "VAGUE", "unknown", ""


def classify(index: Index, header: str, sequence: str, **conf):
    """classify/Classifier.java:341-374 -> the dict Jackson would serialise
    (beans/ClassificationResult.java:59-117, SearchResultEntry.java:53-121), or
    None when the read is dropped (exception logged at LocalClassifier.java:112-114)."""
    r = search(index, sequence, **conf)
    if r is None:
        return None
    hits, _n, _m = r
    entries = []
    if hits:
        top = float(hits[0][1])
        for rank, (d, s) in enumerate(hits):
            if top - float(s) == 0:                   # Classifier.java:360
                doc = index.docs[d]
                entries.append({"docid": d, "rank": rank, "filename": doc.filename,
                                "header": doc.header, "sequence_direction": doc.direction,
                                "taxon_hierarchy": doc.taxonomy, "score": float(s)})
    typ, rank_name, name = label(entries)
    return {"query_header": header, "query": sequence, "result": entries, "type": typ,
            "taxon_rank": rank_name, "taxon_name": name}
