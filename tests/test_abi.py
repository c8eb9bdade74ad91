"""The C-ABI library loads on a CPU-only machine and exports every symbol include/*.h declares.
No compute entry point is called here (there is no GPU); argument validation that happens before any
CUDA call is exercised."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(bsp_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("bsp_index_create", "bsp_index_add_entry", "bsp_index_close", "bsp_classifier_open", "bsp_classify_batch",
                 "bsp_classify_packed", "bsp_classifier_close", "bsp_doc_fields", "bsp_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from biospectra_b200 import _lib
    L = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(set(_lib.EXPORTS)) == declared_symbols()
    assert b"sm_100a" in L.bsp_version()


def test_config_defaults_match_reference():
    # src/biospectra/Configuration.java:37-44
    from biospectra_b200 import _lib
    L = _lib.lib()
    c = _lib.BspConfig()
    L.bsp_config_default(C.byref(c))
    assert (c.kmer_size, c.kmer_skips, c.min_strand_kmer, c.worker_threads, c.index_ram_buffer) == (10, 5, 0, 4, 16)
    assert c.query_algorithm == _lib.PAIRED_PROXIMITY and c.scoring_algorithm == _lib.TFIDF
    assert c.query_term_min_should_match == 0.5


def test_argument_validation_without_gpu(tmp_path):
    # Indexer.java:65-79 / Classifier.java:72-94: IllegalArgumentException cases, checked before any CUDA call
    from biospectra_b200 import _lib
    L = _lib.lib()
    c = _lib.BspConfig()
    L.bsp_config_default(C.byref(c))
    h = C.c_void_p()
    assert L.bsp_index_create(None, b"x", C.byref(h)) == -1
    assert L.bsp_index_create(C.byref(c), None, C.byref(h)) == -1
    c.kmer_size = 0
    assert L.bsp_index_create(C.byref(c), b"x", C.byref(h)) == -1
    assert b"kmerSize" in L.bsp_last_error(None)
    c.kmer_size = 10
    c.worker_threads = 0
    assert L.bsp_index_create(C.byref(c), b"x", C.byref(h)) == -1
    c.worker_threads = 4
    c.kmer_skips = -1
    assert L.bsp_classifier_open(C.byref(c), str(tmp_path).encode(), C.byref(h)) == -1
    c.kmer_skips = 0
    assert L.bsp_classifier_open(C.byref(c), None, C.byref(h)) == -1
    assert L.bsp_classifier_open(C.byref(c), str(tmp_path / "missing").encode(), C.byref(h)) == -1
    assert b"does not exist" in L.bsp_last_error(None)
    # entry points added for the per-read seam and the taxonomy labels: a null handle is an argument error, not a crash
    st = C.c_int32()
    assert L.bsp_classify_one(None, b"ACGT", 4, C.byref(st), None, None, None, None, None) == -1
    assert L.bsp_classifier_set_microbatch(None, 16, 100) == -1
    assert L.bsp_microbatch_stats(None, None, None) == -1
    assert L.bsp_classifier_set_lineages(None, 1, None, None, None) == -1
    assert L.bsp_classify_packed_tax(None, 0, None, None, None, None, None, None, None, None, None, None) == -1
    assert L.bsp_host_classify_local_per_read(None, None, b"", b"", 0, -1) == -1


def test_no_oracle_in_product():
    """the product package must never import or link the oracle"""
    for root, _d, files in os.walk(os.path.join(ROOT, "biospectra_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", "Makefile")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "oracle_lib" not in txt and "spec_oracle" not in txt and "libbsp_oracle" not in txt, f


# ------------------------------------------------------------------ the JNI layer (jni/)
def test_jni_shim_compiles_and_binds_every_native():
    """jni/biospectra_gpu_jni.c parses and type-checks (against jni/stub/jni.h: no JDK in this image) with warnings as
    errors, links against the library's exports, and implements exactly the natives BioSpectraGpu.java declares."""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "jni", "biospectra_gpu_jni.c")
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "jni.o")
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c", "-I", os.path.join(ROOT, "jni", "stub"),
                               "-I", os.path.join(ROOT, "include"), src, "-o", obj])
        # every bsp_* symbol the shim references is exported by the library
        und = subprocess.check_output(["nm", "-u", obj], text=True)
        used = sorted(set(re.findall(r"\b(bsp_[a-z0-9_]+)", und)))
        assert used and "bsp_classify_one" in used and "bsp_index_add_entry" in used
        from biospectra_b200 import _lib
        L = _lib.lib()
        assert not [s for s in used if not hasattr(L, s)]
        defined = subprocess.check_output(["nm", "--defined-only", obj], text=True)
        c_natives = sorted(set(re.findall(r"Java_biospectra_gpu_BioSpectraGpu_(\w+)", defined)))
    java = open(os.path.join(ROOT, "jni", "java", "biospectra", "gpu", "BioSpectraGpu.java")).read()
    j_natives = sorted(set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java)))
    assert c_natives == j_natives, (c_natives, j_natives)
    assert len(j_natives) >= 12


def test_jni_patches_replace_lucene_in_the_two_seam_classes():
    """jni/patches/*.patch (made by jni/patches/make_patches.py against /root/reference/src): after patching, Indexer and
    Classifier call BioSpectraGpu and import nothing from org.apache.lucene."""
    for name, must in (("Indexer.java.patch", ("BioSpectraGpu.indexCreate", "BioSpectraGpu.indexAddEntry", "BioSpectraGpu.indexClose")),
                       ("Classifier.java.patch", ("BioSpectraGpu.classifierOpen", "BioSpectraGpu.classifyOne", "BioSpectraGpu.docFields",
                                                  "BioSpectraGpu.classifierClose"))):
        txt = open(os.path.join(ROOT, "jni", "patches", name)).read()
        added = [l[1:] for l in txt.splitlines() if l.startswith("+") and not l.startswith("+++")]
        removed = [l[1:] for l in txt.splitlines() if l.startswith("-") and not l.startswith("---")]
        for m in must:
            assert any(m in l for l in added), (name, m)
        assert not any("org.apache.lucene" in l for l in added)
        n_lucene_imports = sum(1 for l in removed if l.startswith("import org.apache.lucene"))
        assert n_lucene_imports >= 8, (name, n_lucene_imports)
