"""Host-side mirror of the reference interface (C++ behind the C ABI, no GPU needed): FASTA reader
semantics, configuration.json parsing, label / lowest-common-taxon logic and JSON formatting, each
checked against the literal spec oracle or the reference's documented behaviour."""
import ctypes as C
import gzip
import json
import os
import random

import pytest

import spec_oracle as so


@pytest.fixture(scope="module")
def L():
    from biospectra_b200 import _lib
    return _lib.lib()


def read_fasta_c(L, path, partial=False):
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    need = C.c_int64()
    rc = L.bsp_host_read_fasta(str(path).encode(), buf, cap, C.byref(need))
    recs = buf.value.decode("latin-1").split("\x1e")[:-1]
    recs = [tuple(r.split("\x1f", 1)) for r in recs]
    if partial:
        return recs, (L.bsp_host_last_error().decode() if rc else None)
    if rc:
        raise IOError(L.bsp_host_last_error().decode())
    return recs


with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fasta_reader.json")) as _f:
    FASTA_GOLDEN = json.load(_f)["cases"]


@pytest.mark.parametrize("c", FASTA_GOLDEN, ids=[c["name"] for c in FASTA_GOLDEN])
def test_fasta_reader_matches_the_reference_class_file(L, tmp_path, c):
    """tests/golden/fasta_reader.json: what fasta-utils-1.1.jar!FASTAReader#readNext itself returns (executed in
    oracle/minijvm.py), entry by entry, and whether it ends with an exception."""
    p = tmp_path / (c["name"] + ".fa")
    p.write_bytes(c["text"].encode("latin-1"))
    recs, err = read_fasta_c(L, p, partial=True)
    # bytes >= 0x80 are U+FFFD after the reference's US-ASCII decoding; the host keeps them raw (one byte per char)
    want = [tuple("".join(chr(0xFFFD) if ord(ch) >= 0x80 else ch for ch in x) for x in e) for e in recs]
    assert want == [tuple(e) for e in c["entries"]]
    assert (err is None) == (c["error"] is None), (err, c["error"])


@pytest.mark.parametrize("mmap", ["0", "1"])
def test_fasta_reader_mapped_and_buffered_images_agree(L, tmp_path, monkeypatch, mmap):
    """FASTAReader takes a regular file as a read-only mapping (FileImage, BSP_FASTA_MMAP=1, the default) or reads it into a
    buffer (BSP_FASTA_MMAP=0, also what gzip input and pipes get): every executed-reference case gives the same entries and
    the same error either way; an empty file (nothing to map) yields no entries."""
    monkeypatch.setenv("BSP_FASTA_MMAP", mmap)
    for c in FASTA_GOLDEN:
        p = tmp_path / (c["name"] + ".fa")
        p.write_bytes(c["text"].encode("latin-1"))
        recs, err = read_fasta_c(L, p, partial=True)
        want = [tuple("".join(chr(0xFFFD) if ord(ch) >= 0x80 else ch for ch in x) for x in e) for e in recs]
        assert want == [tuple(e) for e in c["entries"]], c["name"]
        assert (err is None) == (c["error"] is None), (c["name"], err, c["error"])
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    recs, err = read_fasta_c(L, empty, partial=True)
    assert recs == [] and err is None


FASTA_CASES = [
    ">h1 desc\nACGT\nacgtnn\n>h2\nTTTT\n",
    ">h1\r\nAC GT\r\n;comment line\r\nGG\r\n>h2\rTT\r",
    ">only\nACGT",                                   # no trailing newline
    ">a\n  ACGT  \n\n>b\n\tGG\t\n",                  # leading/trailing blanks: only the whole sequence is trimmed
    ">x|y z\nNNNN\nXXXX\n",
]


@pytest.mark.parametrize("txt", FASTA_CASES)
def test_fasta_reader_matches_spec(L, tmp_path, txt):
    # fasta-utils FASTAReader#readNext as restated in SURVEY.md A.1
    p = tmp_path / "x.fa"
    p.write_bytes(txt.encode())
    assert read_fasta_c(L, p) == so.read_fasta(txt)


def test_fasta_reader_gzip_and_errors(L, tmp_path):
    txt = ">g\nACGT\nGGCC\n"
    p = tmp_path / "x.fa.gz"                          # utils/CompressedFileFilter.java:29-32
    with gzip.open(p, "wb") as f:
        f.write(txt.encode())
    assert read_fasta_c(L, p) == [(">g", "ACGTGGCC")]
    bad = tmp_path / "bad.fa"
    bad.write_text("ACGT\n")
    with pytest.raises(IOError):
        read_fasta_c(L, bad)                          # FASTADataErrorException: no '>'
    bad.write_text(">a\n>b\nAC\n")
    with pytest.raises(IOError):
        read_fasta_c(L, bad)                          # header without sequence


def label_c(L, lineages):
    arr = (C.c_char_p * len(lineages))(*[x.encode() for x in lineages])
    buf = C.create_string_buffer(1 << 16)
    assert L.bsp_host_label_json(arr, len(lineages), buf, 1 << 16) == 0, L.bsp_host_last_error()
    d = json.loads(buf.value.decode())
    return d["type"], d["taxon_rank"], d["taxon_name"]


def lineage(*taxa):
    return json.dumps([{"taxid": t, "name": n, "parent": 0, "rank": r} for t, n, r in taxa])


def test_label_logic_matches_spec(L):
    # classify/Classifier.java:263-339 + beans/TaxonTreeDescription.java:70-98
    a = lineage((3, "strain a", "no rank"), (2, "E. coli", "species"), (1, "Escherichia", "genus"))
    b = lineage((4, "strain b", "no rank"), (2, "E. coli", "species"), (1, "Escherichia", "genus"))
    c = lineage((7, "S. enterica", "species"), (6, "Salmonella", "genus"), (5, "Enterobacteriaceae", "family"))
    d = lineage((8, "K. x", "Species"), (5, "Enterobacteriaceae", "family"))
    e = lineage((9, "unranked", ""), (10, "also", "No Rank"))
    cases = [[], [a], [""], [e], [a, b], [a, c], [c, d], [a, ""], ["", a], [a, b, c], [d, c], [e, e], [a, a, a]]
    for cs in cases:
        exp = so.label([{"taxon_hierarchy": x} for x in cs])
        assert label_c(L, cs) == exp, cs
    assert label_c(L, [a, b]) == ("CLASSIFIED", "species", "E. coli")
    assert label_c(L, [c, d]) == ("CLASSIFIED", "family", "Enterobacteriaceae")
    assert label_c(L, [a, c]) == ("VAGUE", "unknown", "")
    assert label_c(L, []) == ("UNKNOWN", "unknown", "")
    assert label_c(L, [""]) == ("CLASSIFIED", "unknown", "")


def test_label_random_lineages(L):
    rng = random.Random(3)
    ranks = ["species", "genus", "family", "no rank", "", "order"]
    for _ in range(200):
        n = rng.randint(1, 4)
        ls = []
        for _j in range(n):
            if rng.random() < 0.1:
                ls.append("")
                continue
            ls.append(lineage(*[(rng.randint(1, 6), "n%d" % rng.randint(0, 9), rng.choice(ranks)) for _t in range(rng.randint(0, 4))]))
        assert label_c(L, ls) == so.label([{"taxon_hierarchy": x} for x in ls]), ls


def test_parse_configuration(L):
    from biospectra_b200 import _lib
    cfg = _lib.BspConfig()
    path = C.create_string_buffer(256)
    shipped = json.dumps({                            # the keys of the reference's shipped configuration.json
        "index_path": "/tmp/idx", "kmer_size": 10, "kmer_skips": 0, "min_strand_kmer": False,
        "query_term_min_should_match": 0.5, "worker_threads": 4, "index_ram_buffer": 16,
        "query_algorithm": "PAIRED_PROXIMITY", "scoring_algorithm": "tfidf"})
    assert L.bsp_host_parse_config(shipped.encode(), C.byref(cfg), path, 256) == 0
    assert (cfg.kmer_size, cfg.kmer_skips, cfg.min_strand_kmer, cfg.query_algorithm, cfg.scoring_algorithm) == (10, 0, 0, 2, 0)
    assert path.value == b"/tmp/idx" and cfg.query_term_min_should_match == 0.5
    # defaults for missing keys (Configuration.java:37-44); "bm25" in any case selects BM25 (:169-179)
    assert L.bsp_host_parse_config(b'{"scoring_algorithm":"BM25","query_algorithm":"NAIVE_KMER"}', C.byref(cfg), path, 256) == 0
    assert (cfg.kmer_size, cfg.kmer_skips, cfg.scoring_algorithm, cfg.query_algorithm) == (10, 5, 1, 0)
    # unknown keys are rejected like Jackson's FAIL_ON_UNKNOWN_PROPERTIES (utils/JsonSerializer.java:33-35)
    assert L.bsp_host_parse_config(b'{"kmer_size":10,"gpu_device":1}', C.byref(cfg), path, 256) != 0
    assert L.bsp_host_parse_config(b'', C.byref(cfg), path, 256) != 0


def test_python_configuration_mirror():
    from biospectra_b200 import Configuration
    c = Configuration.from_json('{"kmer_size": 12, "scoring_algorithm": "bm25", "query_algorithm": "CHAIN_PROXIMITY"}')
    s = c.to_struct()
    assert (s.kmer_size, s.kmer_skips, s.scoring_algorithm, s.query_algorithm) == (12, 5, 1, 1)
    with pytest.raises(ValueError):
        Configuration.from_json('{"devices": 8}')
    assert json.loads(c.to_json())["query_algorithm"] == "CHAIN_PROXIMITY"


@pytest.mark.parametrize("x,s", [(1.0, "1.0"), (0.5, "0.5"), (0.0, "0.0"), (100.0, "100.0"), (1e7, "1.0E7"), (1e-3, "0.001"),
                                 (1e-4, "1.0E-4"), (0.9399199485778809, "0.9399199485778809"), (123456.789, "123456.789"),
                                 (9999999.0, "9999999.0"), (0.08867590874433517, "0.08867590874433517")])
def test_java_double_formatting(L, x, s):
    # Jackson writes Double.toString(score) (beans/SearchResultEntry.java "score")
    buf = C.create_string_buffer(64)
    assert L.bsp_host_double_to_string(x, buf, 64) == 0
    assert buf.value.decode() == s
    assert float(buf.value) == x


def _read_both(L, path, stripe):
    out = []
    for fn, args in ((L.bsp_host_read_fasta, ()), (L.bsp_host_read_fasta_striped, (C.c_int64(stripe),))):
        cap = 1 << 22
        buf = C.create_string_buffer(cap)
        need = C.c_int64()
        rc = fn(str(path).encode(), *args, buf, cap, C.byref(need))
        out.append((rc, buf.value, L.bsp_host_last_error() if rc else b""))
    return out


@pytest.mark.parametrize("c", FASTA_GOLDEN, ids=[c["name"] for c in FASTA_GOLDEN])
@pytest.mark.parametrize("stripe", [1, 7, 64])
def test_striped_fasta_reader_on_the_reference_cases(L, tmp_path, c, stripe):
    """The batch driver parses a read file in stripes cut at entry starts (several parser threads): whatever the stripe
    size, the entries, their order, the return code and the error text are those of the one-pass reader -- on the
    executed-reference cases (CR / CRLF line ends, ';' lines, headers without a sequence, malformed headers)."""
    p = tmp_path / (c["name"] + ".fa")
    p.write_bytes(c["text"].encode("latin-1"))
    whole, striped = _read_both(L, p, stripe)
    assert whole == striped


def test_striped_fasta_reader_random_files(L, tmp_path):
    rng = random.Random(12)
    eols = ["\n", "\r\n", "\r"]
    for case in range(60):
        parts = []
        for e in range(rng.randint(1, 40)):
            eol = rng.choice(eols)
            parts.append(">" + "".join(rng.choice("abc|> \t;1") for _ in range(rng.randint(1, 12))).lstrip(" \t") + "x" + eol)
            for _line in range(rng.randint(0 if rng.random() < 0.05 else 1, 4)):
                if rng.random() < 0.1:
                    parts.append(";" + "c" * rng.randint(0, 5) + eol)
                body = "".join(rng.choice("ACGTNacgt >;") for _ in range(rng.randint(0, 30)))
                if body.startswith(">"):
                    body = "A" + body
                parts.append(body + eol)
        text = "".join(parts)
        if rng.random() < 0.3:
            text = text.rstrip("\r\n")
        p = tmp_path / ("r%d.fa" % case)
        p.write_bytes(text.encode("latin-1"))
        for stripe in (1, 13, 100, 1 << 20):
            whole, striped = _read_both(L, p, stripe)
            assert whole == striped, (case, stripe, text)
