"""ctypes binding of oracle/libbsp_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see bsp_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class QueryConf(C.Structure):
    _fields_ = [("kmer_skips", C.c_int32), ("query_algorithm", C.c_int32),
                ("scoring_algorithm", C.c_int32), ("threads", C.c_int32),
                ("query_term_min_should_match", C.c_double)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libbsp_oracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.ora_index_new.restype = C.c_void_p
    L.ora_index_new.argtypes = [C.c_int, C.c_int]
    L.ora_index_add_entry.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.ora_index_finish.argtypes = [C.c_void_p]
    L.ora_index_free.argtypes = [C.c_void_p]
    L.ora_index_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
    L.ora_doc_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ora_term_postings.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int64),
                                    C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int64)),
                                    C.POINTER(C.POINTER(C.c_int32))]
    L.ora_term_at.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64)]
    L.ora_tokenize.restype = C.c_int64
    L.ora_tokenize.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    L.ora_sloppy_freq.restype = C.c_float
    L.ora_sloppy_freq.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ora_float_to_byte315.argtypes = [C.c_float]
    L.ora_byte315_to_float.restype = C.c_float
    L.ora_byte315_to_float.argtypes = [C.c_int]
    L.ora_norm_byte.argtypes = [C.c_int64]
    L.ora_idf_tfidf.restype = C.c_float
    L.ora_idf_tfidf.argtypes = [C.c_int64, C.c_int64]
    L.ora_idf_bm25.restype = C.c_float
    L.ora_idf_bm25.argtypes = [C.c_int64, C.c_int64]
    L.ora_classify_batch.argtypes = [C.c_void_p, C.POINTER(QueryConf), C.c_int32, C.POINTER(C.c_char_p),
                                     C.c_void_p] + [C.c_void_p] * 8
    L.ora_max_threads.restype = C.c_int
    L.ora_index_dense_df.argtypes = [C.c_void_p, C.c_void_p]
    L.ora_index_set_global.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.ora_index_sum_ttf.restype = C.c_int64
    L.ora_index_sum_ttf.argtypes = [C.c_void_p]
    _LIB = L
    return L


def tokenize(seq: bytes, k: int, skips: int, min_strand: bool = False):
    L = lib()
    cap = max(len(seq), 1)
    keys = np.zeros(cap, np.uint64)
    offs = np.zeros(cap, np.int32)
    n = L.ora_tokenize(seq, len(seq), k, skips, int(min_strand), keys.ctypes.data, offs.ctypes.data, cap)
    return keys[:n].copy(), offs[:n].copy()


def sloppy_freq(p0, p1, slop, same_term=False) -> np.float32:
    a = np.asarray(p0, np.int32)
    b = np.asarray(p1, np.int32)
    return np.float32(lib().ora_sloppy_freq(a.ctypes.data, len(a), b.ctypes.data, len(b), slop, int(same_term)))


class OracleIndex:
    def __init__(self, k: int, min_strand: bool = False):
        self.L = lib()
        self.k, self.min_strand = k, bool(min_strand)
        self.h = self.L.ora_index_new(k, int(min_strand))
        if not self.h:
            raise ValueError("bad k")
        self.n_entries = 0

    def add_entry(self, seq: bytes) -> int:
        self.n_entries += 1
        return self.L.ora_index_add_entry(self.h, seq, len(seq))

    def finish(self):
        self.L.ora_index_finish(self.h)
        return self

    def stats(self):
        v = [C.c_int64() for _ in range(4)]
        self.L.ora_index_stats(self.h, *[C.byref(x) for x in v])
        return dict(zip(("n_docs", "n_terms", "n_postings", "n_positions"), [x.value for x in v]))

    def doc_info(self, d):
        a, b = C.c_int32(), C.c_int32()
        if self.L.ora_doc_info(self.h, d, C.byref(a), C.byref(b)):
            raise IndexError(d)
        return a.value, b.value

    def term_at(self, i):
        k = C.c_uint64()
        if self.L.ora_term_at(self.h, i, C.byref(k)):
            raise IndexError(i)
        return k.value

    def postings(self, key: int):
        """-> (docs int32[df], tfs int32[df], positions list-of-arrays)"""
        df = C.c_int64()
        docs = C.POINTER(C.c_int32)()
        poff = C.POINTER(C.c_int64)()
        pos = C.POINTER(C.c_int32)()
        self.L.ora_term_postings(self.h, key, C.byref(df), C.byref(docs), C.byref(poff), C.byref(pos))
        n = df.value
        if n == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), []
        d = np.ctypeslib.as_array(docs, (n,)).copy()
        po = np.ctypeslib.as_array(poff, (n + 1,)).copy()
        a, b = int(po[0]), int(po[-1])
        addr = C.addressof(pos.contents) + 4 * a
        flat = np.ctypeslib.as_array((C.c_int32 * (b - a)).from_address(addr)).copy()
        po -= a
        return d, np.diff(po).astype(np.int32), [flat[po[i]:po[i + 1]] for i in range(n)]

    def classify(self, seqs, *, skips=0, algorithm=2, scoring=0, msm=0.5, threads=1):
        n = len(seqs)
        arr = (C.c_char_p * n)(*seqs)
        lens = np.array([len(s) for s in seqs], np.int32)
        out = {
            "status": np.zeros(n, np.int32), "label": np.zeros(n, np.int32), "n_hits": np.zeros(n, np.int32),
            "n_top": np.zeros(n, np.int32), "top_doc": np.full((n, 10), -1, np.int32),
            "top_score": np.zeros((n, 10), np.float32), "n_clauses": np.zeros(n, np.int32),
            "msm": np.zeros(n, np.int32),
        }
        qc = QueryConf(skips, algorithm, scoring, threads, msm)
        rc = self.L.ora_classify_batch(self.h, C.byref(qc), n, arr, lens.ctypes.data,
                                       *[out[k].ctypes.data for k in ("status", "label", "n_hits", "n_top",
                                                                      "top_doc", "top_score", "n_clauses", "msm")])
        if rc:
            raise RuntimeError("ora_classify_batch rc=%d" % rc)
        return out

    # ---- doc-partitioned checks: collection-wide statistics
    def dense_df(self):
        out = np.zeros(1 << (2 * self.k), np.uint32)
        if self.L.ora_index_dense_df(self.h, out.ctypes.data):
            raise ValueError("dense df needs k <= 12")
        return out

    def sum_ttf(self):
        return int(self.L.ora_index_sum_ttf(self.h))

    def set_global(self, max_doc, sum_ttf, dense_df):
        d = np.ascontiguousarray(dense_df, np.uint32)
        if self.L.ora_index_set_global(self.h, int(max_doc), int(sum_ttf), d.ctypes.data):
            raise ValueError("set_global failed")

    def close(self):
        if self.h:
            self.L.ora_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
