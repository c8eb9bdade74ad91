"""ctypes binding of libbiospectra_gpu.so (the C ABI declared in include/biospectra_gpu.h).

There is NO CPU fallback: importing this module fails loudly when the CUDA library has
not been built (run `python -c "import __graft_entry__ as g; g.build()"` or
`make -C biospectra_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSP_LIB: another build of the same library (A/B runs of build-time variants)
LIB_PATH = os.environ.get("BSP_LIB") or os.path.join(_HERE, "libbiospectra_gpu.so")

BSP_TOPN = 10
NAIVE_KMER, CHAIN_PROXIMITY, PAIRED_PROXIMITY = 0, 1, 2
TFIDF, BM25 = 0, 1
READ_OK, READ_DROPPED, READ_ERROR = 0, 1, 2
UNKNOWN, VAGUE, CLASSIFIED = 0, 1, 2
LABEL_NAMES = ("UNKNOWN", "VAGUE", "CLASSIFIED")
ROUTE_DENSE, ROUTE_POSITIONAL, ROUTE_FILTER = 1, 2, 3


class BspConfig(C.Structure):
    _fields_ = [
        ("kmer_size", C.c_int32), ("kmer_skips", C.c_int32), ("min_strand_kmer", C.c_int32),
        ("query_algorithm", C.c_int32), ("scoring_algorithm", C.c_int32), ("worker_threads", C.c_int32),
        ("index_ram_buffer", C.c_int32), ("device", C.c_int32), ("query_term_min_should_match", C.c_double),
        ("part_rank", C.c_int32), ("part_count", C.c_int32), ("dense_tables", C.c_int32), ("positional", C.c_int32),
        ("full_topn", C.c_int32),
    ]


class BspError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bsp error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


# every symbol include/biospectra_gpu.h declares
EXPORTS = [
    "bsp_config_default", "bsp_last_error", "bsp_version", "bsp_index_create", "bsp_index_add_entry",
    "bsp_index_add_entry_device", "bsp_index_close", "bsp_index_close_to_classifier", "bsp_classifier_open",
    "bsp_classify_batch", "bsp_classify_packed", "bsp_classify_device", "bsp_classify_one",
    "bsp_classifier_set_microbatch", "bsp_microbatch_stats", "bsp_classifier_set_lineages", "bsp_classify_packed_tax", "bsp_doc_fields", "bsp_classifier_close",
    "bsp_index_stats", "bsp_doc_info", "bsp_term_postings", "bsp_tokenize", "bsp_last_routes", "bsp_last_counters",
    "bsp_last_timings", "bsp_last_gap_join", "bsp_memory_info", "bsp_partition_df_buffer", "bsp_partition_local_stats",
    "bsp_partition_set_global", "bsp_merge_topk_device", "bsp_read_cost", "bsp_classifier_set_query", "bsp_host_classify_file", "bsp_index_layout", "bsp_pack_topk_device", "bsp_merge_topk_records_device",
    "bsp_host_index", "bsp_host_classify_local", "bsp_host_classify_local_per_read", "bsp_host_last_error", "bsp_host_label_json", "bsp_host_read_fasta", "bsp_host_read_fasta_striped", "bsp_index_add_fasta",
    "bsp_host_double_to_string", "bsp_host_parse_config",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libbiospectra_gpu.so is not built (%s); there is no CPU fallback -- build it with "
                          "`make -C biospectra_b200/csrc`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    cfgp = C.POINTER(BspConfig)
    L.bsp_config_default.argtypes = [cfgp]
    L.bsp_config_default.restype = None
    L.bsp_last_error.argtypes = [vp]
    L.bsp_last_error.restype = C.c_char_p
    L.bsp_version.restype = C.c_char_p
    L.bsp_classifier_set_query.argtypes = [vp, cfgp]
    L.bsp_pack_topk_device.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.bsp_merge_topk_records_device.argtypes = [vp, C.c_int32, C.c_int32] + [vp] * 8
    L.bsp_index_layout.argtypes = [vp, i64p]
    L.bsp_host_classify_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    L.bsp_index_create.argtypes = [cfgp, C.c_char_p, C.POINTER(vp)]
    L.bsp_index_add_entry.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]
    L.bsp_index_add_entry_device.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, vp, C.c_int64]
    L.bsp_index_close.argtypes = [vp]
    L.bsp_index_close_to_classifier.argtypes = [vp, cfgp, C.POINTER(vp)]
    L.bsp_classifier_open.argtypes = [cfgp, C.c_char_p, C.POINTER(vp)]
    L.bsp_classify_batch.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p), vp, vp, vp, vp, vp, vp, vp]
    L.bsp_classify_packed.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bsp_classifier_set_lineages.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.bsp_classify_packed_tax.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bsp_classify_one.argtypes = [vp, C.c_char_p, C.c_int32, i32p, i32p, i32p, i32p, vp, vp]
    L.bsp_classifier_set_microbatch.argtypes = [vp, C.c_int32, C.c_int64]
    L.bsp_microbatch_stats.argtypes = [vp, i64p, i64p]
    L.bsp_classify_device.argtypes = [vp, C.c_int32, vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.bsp_doc_fields.argtypes = [vp, C.c_int32] + [C.POINTER(C.c_char_p)] * 4
    L.bsp_classifier_close.argtypes = [vp]
    L.bsp_index_stats.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.bsp_doc_info.argtypes = [vp, C.c_int32, i32p, i32p]
    L.bsp_term_postings.argtypes = [vp, C.c_uint64, i64p, i64p, vp, vp, vp]
    L.bsp_tokenize.argtypes = [vp, C.c_char_p, C.c_int32, vp, vp, C.c_int64]
    L.bsp_tokenize.restype = C.c_int64
    L.bsp_last_routes.argtypes = [vp, C.c_int32, vp]
    L.bsp_last_counters.argtypes = [vp, vp]
    L.bsp_last_timings.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bsp_last_gap_join.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.bsp_memory_info.argtypes = [vp, i64p, i64p, i64p]
    L.bsp_partition_df_buffer.argtypes = [vp, C.POINTER(vp), i64p]
    L.bsp_partition_local_stats.argtypes = [vp, i64p, i64p]
    L.bsp_partition_set_global.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.bsp_merge_topk_device.argtypes = [vp, C.c_int32, C.c_int32] + [vp] * 10
    L.bsp_read_cost.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.bsp_host_index.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    L.bsp_host_classify_local.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    L.bsp_host_classify_local_per_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
    L.bsp_host_last_error.restype = C.c_char_p
    L.bsp_host_label_json.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_char_p, C.c_int32]
    L.bsp_host_read_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, i64p]
    L.bsp_index_add_fasta.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.bsp_host_double_to_string.argtypes = [C.c_double, C.c_char_p, C.c_int32]
    L.bsp_host_parse_config.argtypes = [C.c_char_p, cfgp, C.c_char_p, C.c_int32]
    _lib = L
    return L


def check(rc, handle=None):
    if rc < 0:
        msg = lib().bsp_last_error(handle)
        raise BspError(rc, (msg or b"").decode("utf-8", "replace"))
    return rc
