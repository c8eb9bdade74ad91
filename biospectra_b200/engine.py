"""Thin Python mirror of the two reference classes the C ABI replaces:
biospectra.index.Indexer (index/Indexer.java) and biospectra.classify.Classifier
(classify/Classifier.java).  All work happens in libbiospectra_gpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .config import Configuration


def _b(s):
    if s is None:
        return None
    return s if isinstance(s, bytes) else s.encode("utf-8")


class Indexer:
    """Indexer(Configuration); index entries; close() builds + writes the index."""

    def __init__(self, conf: Configuration):
        if conf is None:
            raise ValueError("conf is null")
        self.conf = conf
        self.L = _lib.lib()
        h = C.c_void_p()
        cs = conf.to_struct()
        _lib.check(self.L.bsp_index_create(C.byref(cs), _b(conf.index_path) if conf.index_path is not None else None,
                                           C.byref(h)))
        self.h = h

    def add_entry(self, filename, header, taxonomy, sequence):
        seq = _b(sequence)
        _lib.check(self.L.bsp_index_add_entry(self.h, _b(filename), _b(header), _b(taxonomy or ""), seq, len(seq)), self.h)

    def add_fasta(self, fasta_path, taxd_path=None):
        """Indexer.index(File fastaDoc, File taxonDoc) (Indexer.java:150-236) through bsp_index_add_fasta."""
        rc = self.L.bsp_index_add_fasta(self.h, _b(str(fasta_path)), _b(str(taxd_path)) if taxd_path else None)
        if rc:
            raise IOError(self.L.bsp_host_last_error().decode("utf-8", "replace"))

    def add_entry_device(self, filename, header, taxonomy, dev_ptr: int, length: int):
        _lib.check(self.L.bsp_index_add_entry_device(self.h, _b(filename), _b(header), _b(taxonomy or ""),
                                                     C.c_void_p(dev_ptr), length), self.h)

    def close(self):
        if self.h:
            h, self.h = self.h, None
            _lib.check(self.L.bsp_index_close(h))

    def close_to_classifier(self, query_conf: Configuration) -> "Classifier":
        out = C.c_void_p()
        cs = query_conf.to_struct()
        h, self.h = self.h, None
        _lib.check(self.L.bsp_index_close_to_classifier(h, C.byref(cs), C.byref(out)))
        return Classifier(query_conf, _handle=out)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Classifier:
    """Classifier(Configuration) + batched classify."""

    def __init__(self, conf: Configuration, _handle=None):
        if conf is None:
            raise ValueError("conf is null")
        self.conf = conf
        self.L = _lib.lib()
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            cs = conf.to_struct()
            _lib.check(self.L.bsp_classifier_open(C.byref(cs), _b(conf.index_path) if conf.index_path is not None else None,
                                                  C.byref(h)))
            self.h = h

    # ------------------------------------------------------------- classification
    @staticmethod
    def pack_reads(seqs):
        lens = np.fromiter((len(s) for s in seqs), np.int64, len(seqs))
        off = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        data = b"".join(_b(s) for s in seqs)
        return np.frombuffer(data, np.uint8), off

    @staticmethod
    def alloc_outputs(n: int, pinned: bool = False):
        """result arrays of bsp_classify_packed; pinned=True page-locks them (torch) so that the copy-out of one
        sub-batch overlaps the kernels of the next"""
        shapes = {"status": ((n,), np.int32), "label": ((n,), np.int32), "n_hits": ((n,), np.int32), "n_top": ((n,), np.int32),
                  "top_doc": ((n, 10), np.int32), "top_score": ((n, 10), np.float32)}
        if not pinned:
            return {k: np.zeros(sh, dt) for k, (sh, dt) in shapes.items()}
        import torch
        keep = {k: torch.zeros(sh, dtype=torch.int32 if dt == np.int32 else torch.float32, pin_memory=True)
                for k, (sh, dt) in shapes.items()}
        out = {k: v.numpy() for k, v in keep.items()}
        out["_pinned"] = keep
        return out

    def classify_packed(self, seq_bytes: np.ndarray, offsets: np.ndarray, out=None):
        n = len(offsets) - 1
        if out is None:
            out = self.alloc_outputs(n)
        seq_bytes = np.ascontiguousarray(seq_bytes, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        _lib.check(self.L.bsp_classify_packed(self.h, n, seq_bytes.ctypes.data if seq_bytes.size else None, offsets.ctypes.data,
                                              *[out[k].ctypes.data for k in ("status", "label", "n_hits", "n_top", "top_doc",
                                                                             "top_score")]), self.h)
        return out

    def classify_batch(self, seqs):
        if any((s is None or len(s) == 0) for s in seqs):
            raise ValueError("sequence is null or empty")      # Classifier.java:342-344
        data, off = self.pack_reads(seqs)
        return self.classify_packed(data, off)

    def set_lineages(self, lineage_off, taxids, flags):
        """interned ranked lineages per document (leaf -> root taxids of the entries with a rank); flags bit0 = has a
        taxonomy string, bit1 = malformed.  Enables classify_packed_tax (Classifier.java:263-339 on the device)."""
        off = np.ascontiguousarray(lineage_off, np.int64)
        tax = np.ascontiguousarray(taxids, np.int32)
        fl = np.ascontiguousarray(flags, np.uint8)
        _lib.check(self.L.bsp_classifier_set_lineages(self.h, len(fl), off.ctypes.data, tax.ctypes.data if tax.size else None,
                                                      fl.ctypes.data), self.h)

    def classify_packed_tax(self, seq_bytes: np.ndarray, offsets: np.ndarray):
        n = len(offsets) - 1
        out = self.alloc_outputs(n)
        out["tax_label"] = np.zeros(n, np.int32)
        out["tax_index"] = np.full(n, -1, np.int32)
        seq_bytes = np.ascontiguousarray(seq_bytes, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        _lib.check(self.L.bsp_classify_packed_tax(self.h, n, seq_bytes.ctypes.data if seq_bytes.size else None, offsets.ctypes.data,
                                                  *[out[k].ctypes.data for k in ("status", "label", "n_hits", "n_top", "top_doc",
                                                                                 "top_score", "tax_label", "tax_index")]), self.h)
        return out

    def classify_one(self, seq: bytes):
        """One read, blocking; safe to call from many threads at once (the GIL is released inside the call): concurrent
        calls are coalesced into GPU micro-batches.  Mirrors Classifier.classify(header, sequence), Classifier.java:341-366."""
        if seq is None or len(seq) == 0:
            raise ValueError("sequence is null or empty")      # Classifier.java:342-344
        st, label, nh, nt = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        doc = (C.c_int32 * 10)()
        score = (C.c_float * 10)()
        _lib.check(self.L.bsp_classify_one(self.h, seq, len(seq), C.byref(st), C.byref(label), C.byref(nh), C.byref(nt), doc, score), self.h)
        return {"status": st.value, "label": label.value, "n_hits": nh.value, "n_top": nt.value,
                "top_doc": np.array(doc[:], np.int32), "top_score": np.array(score[:], np.float32)}

    def set_microbatch(self, max_reads: int, max_wait_us: int):
        _lib.check(self.L.bsp_classifier_set_microbatch(self.h, max_reads, max_wait_us), self.h)

    def microbatch_stats(self):
        b, r = C.c_int64(), C.c_int64()
        _lib.check(self.L.bsp_microbatch_stats(self.h, C.byref(b), C.byref(r)), self.h)
        return {"batches": b.value, "reads": r.value}

    def classify_device(self, n, d_bytes, d_offsets, total_bytes, d_status, d_label, d_n_hits, d_n_top, d_top_doc, d_top_score):
        vp = C.c_void_p
        _lib.check(self.L.bsp_classify_device(self.h, n, vp(d_bytes), vp(d_offsets), total_bytes, 0, vp(d_status), vp(d_label),
                                              vp(d_n_hits), vp(d_n_top), vp(d_top_doc), vp(d_top_score)), self.h)

    # ------------------------------------------------------------- introspection
    def set_query(self, conf: Configuration):
        """new query-side settings on the open index (bsp_classifier_set_query): the reference constructs another
        Classifier(Configuration) on the same index directory for this"""
        cs = conf.to_struct()
        _lib.check(self.L.bsp_classifier_set_query(self.h, C.byref(cs)), self.h)
        self.conf = conf

    def classify_file(self, input_fasta, result_path, summary_path=None, worker_threads=4):
        """LocalClassifier.classify(inputFasta, classifyOutput, summaryOutput) on this open index (bsp_host_classify_file)"""
        rc = self.L.bsp_host_classify_file(self.h, _b(str(input_fasta)), _b(str(result_path)),
                                           _b(str(summary_path)) if summary_path else None, worker_threads)
        if rc:
            raise IOError(self.L.bsp_host_last_error().decode("utf-8", "replace"))

    def doc_fields(self, doc: int):
        p = [C.c_char_p() for _ in range(4)]
        _lib.check(self.L.bsp_doc_fields(self.h, doc, *[C.byref(x) for x in p]), self.h)
        return tuple((x.value or b"").decode("utf-8", "replace") for x in p)

    def stats(self):
        v = [C.c_int64() for _ in range(4)]
        _lib.check(self.L.bsp_index_stats(self.h, *[C.byref(x) for x in v]), self.h)
        return dict(zip(("n_docs", "n_terms", "n_postings", "n_positions"), [x.value for x in v]))

    def doc_info(self, doc: int):
        a, b = C.c_int32(), C.c_int32()
        _lib.check(self.L.bsp_doc_info(self.h, doc, C.byref(a), C.byref(b)), self.h)
        return a.value, b.value

    def postings(self, key: int):
        df, ttf = C.c_int64(), C.c_int64()
        _lib.check(self.L.bsp_term_postings(self.h, key, C.byref(df), C.byref(ttf), None, None, None), self.h)
        docs = np.zeros(df.value, np.int32)
        tfs = np.zeros(df.value, np.int32)
        pos = np.zeros(ttf.value, np.int32)
        if df.value:
            _lib.check(self.L.bsp_term_postings(self.h, key, C.byref(df), C.byref(ttf), docs.ctypes.data, tfs.ctypes.data,
                                                pos.ctypes.data), self.h)
        return docs, tfs, pos

    def tokenize(self, seq):
        s = _b(seq)
        cap = max(len(s), 1)
        keys = np.zeros(cap, np.uint64)
        offs = np.zeros(cap, np.int32)
        n = self.L.bsp_tokenize(self.h, s, len(s), keys.ctypes.data, offs.ctypes.data, cap)
        _lib.check(int(n), self.h)
        return keys[:n].copy(), offs[:n].copy()

    def last_routes(self, n):
        r = np.zeros(n, np.int32)
        _lib.check(self.L.bsp_last_routes(self.h, n, r.ctypes.data), self.h)
        return r

    def last_counters(self):
        c = np.zeros(8, np.int64)
        _lib.check(self.L.bsp_last_counters(self.h, c.ctypes.data), self.h)
        return dict(zip(("reads", "dense_reads", "positional_reads", "filter_reads", "launches", "filter_candidates",
                         "filter_exception_cells", "filter_overflow_reads"), [int(x) for x in c]))

    def last_timings(self):
        a, b = C.c_double(), C.c_double()
        _lib.check(self.L.bsp_last_timings(self.h, C.byref(a), C.byref(b)), self.h)
        return {"score_kernel_ms": a.value, "pipeline_ms": b.value}

    def last_gap_join(self):
        c = np.zeros(4, np.int64)
        ms, ms_v, cand = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(self.L.bsp_last_gap_join(self.h, c.ctypes.data, C.byref(ms), C.byref(ms_v), C.byref(cand)), self.h)
        return dict(zip(("clauses", "chunks", "matches", "position_bytes"), [int(x) for x in c]), join_kernel_ms=ms.value,
                    verify_kernel_ms=ms_v.value, candidates=cand.value)

    def memory_info(self):
        v = [C.c_int64() for _ in range(3)]
        _lib.check(self.L.bsp_memory_info(self.h, *[C.byref(x) for x in v]), self.h)
        return dict(zip(("positional_bytes", "table_bytes", "staged_bytes"), [x.value for x in v]))

    def layout(self):
        out = (C.c_int64 * 8)()
        _lib.check(self.L.bsp_index_layout(self.h, out), self.h)
        keys = ("row_bytes", "d_pad", "segments", "max_seg_docs", "tables", "positional", "has_pair", "hash_dictionary")
        return dict(zip(keys, [int(x) for x in out]))

    def read_cost(self, seq_bytes: np.ndarray, offsets: np.ndarray):
        """Cost model of SURVEY.md 8(d) evaluated on the actual index (see bsp_read_cost)."""
        out = np.zeros(8, np.int64)
        n = len(offsets) - 1
        seq_bytes = np.ascontiguousarray(seq_bytes, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        _lib.check(self.L.bsp_read_cost(self.h, n, seq_bytes.ctypes.data, offsets.ctypes.data, out.ctypes.data), self.h)
        return dict(zip(("reads", "tokens", "unique_terms", "sum_df", "sum_ttf", "bytes_csr", "bytes_dense", "clauses"),
                        [int(x) for x in out]))

    # ------------------------------------------------------------- partition plumbing
    def partition_df_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self.L.bsp_partition_df_buffer(self.h, C.byref(p), C.byref(n)), self.h)
        return p.value, n.value

    def partition_local_stats(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self.L.bsp_partition_local_stats(self.h, C.byref(a), C.byref(b)), self.h)
        return a.value, b.value

    def partition_set_global(self, doc_base, global_docs, global_sum_ttf):
        _lib.check(self.L.bsp_partition_set_global(self.h, doc_base, global_docs, global_sum_ttf), self.h)

    def merge_topk_device(self, n, parts, p_ntop, p_doc, p_score, p_status, status, label, n_hits, n_top, top_doc, top_score):
        vp = C.c_void_p
        _lib.check(self.L.bsp_merge_topk_device(self.h, n, parts, vp(p_ntop), vp(p_doc), vp(p_score), vp(p_status), vp(status),
                                                vp(label), vp(n_hits), vp(n_top), vp(top_doc), vp(top_score)), self.h)

    def pack_topk_device(self, n, n_top, top_doc, top_score, records):
        vp = C.c_void_p
        _lib.check(self.L.bsp_pack_topk_device(self.h, n, vp(n_top), vp(top_doc), vp(top_score), vp(records)), self.h)

    def merge_topk_records_device(self, n, parts, records, p_status, status, label, n_hits, n_top, top_doc, top_score):
        vp = C.c_void_p
        _lib.check(self.L.bsp_merge_topk_records_device(self.h, n, parts, vp(records), vp(p_status), vp(status), vp(label),
                                                        vp(n_hits), vp(n_top), vp(top_doc), vp(top_score)), self.h)

    def close(self):
        if self.h:
            h, self.h = self.h, None
            _lib.check(self.L.bsp_classifier_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
