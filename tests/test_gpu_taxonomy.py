"""Taxonomy-aware labels on the device: taxonomy_label_kernel restates
Classifier.makeClassificationResult (classify/Classifier.java:263-339) on interned lineages.  Checked against the spec
oracle's `label` (literal restatement of the Java) applied to the same hit lists, on a corpus built to produce every
branch: single hits, ties inside a species / a genus / across genera, hits without taxonomy, unranked-only lineages."""
import json
import random

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

TYPE = {"UNKNOWN": 0, "VAGUE": 1, "CLASSIFIED": 2}


def _lineage(strain, species, genus, ranks=("no rank", "species", "genus")):
    return json.dumps([{"taxid": 1000 + strain, "name": "strain %d" % strain, "parent": 100 + species, "rank": ranks[0]},
                       {"taxid": 100 + species, "name": "species %d" % species, "parent": 10 + genus, "rank": ranks[1]},
                       {"taxid": 10 + genus, "name": "genus %d" % genus, "parent": 1, "rank": ranks[2]},
                       {"taxid": 1, "name": "root", "parent": 0, "rank": "no rank"}])


def test_device_taxonomy_labels_match_spec_oracle():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import spec_oracle
    from biospectra_b200 import Configuration, Indexer
    rng = random.Random(17)
    core = [common.rand_seq(rng, 400) for _ in range(4)]
    ents = []      # (sequence, taxonomy)
    ents.append((core[0], _lineage(0, 0, 0)))
    ents.append((core[0], _lineage(1, 0, 0)))                      # same species as 0: ties classify at the species
    ents.append((core[1], _lineage(2, 1, 0)))
    ents.append((core[1], _lineage(3, 2, 0)))                      # same genus only
    ents.append((core[2], _lineage(4, 3, 1)))
    ents.append((core[2], _lineage(5, 4, 2)))                      # nothing ranked in common: VAGUE
    ents.append((core[3], _lineage(6, 5, 3)))
    ents.append((core[3], ""))                                     # a hit without taxonomy: VAGUE, quick fail
    ents.append((common.rand_seq(rng, 400), _lineage(7, 6, 4, ranks=("no rank", "", "No Rank"))))   # nothing classifiable
    ents.append((common.rand_seq(rng, 400), ""))                   # single hit without taxonomy
    ents.append((common.rand_seq(rng, 400), _lineage(8, 7, 5)))    # plain single hits
    k = 8
    ix = Indexer(Configuration(index_path="", kmer_size=k))
    for i, (s, t) in enumerate(ents):
        ix.add_entry("f%d.fa" % i, "h%d" % i, t, s)
    g = ix.close_to_classifier(Configuration(index_path="", kmer_size=k, kmer_skips=0, query_algorithm="PAIRED_PROXIMITY",
                                             query_term_min_should_match=0.5))
    try:
        # documents: forward 2e, reverse 2e+1, both carry the entry's taxonomy
        doc_tax = [t for (_s, t) in ents for _ in range(2)]
        off, tax, flags = [0], [], []
        for t in doc_tax:
            flags.append(1 if t else 0)
            if t:
                tax += [int(x["taxid"]) for x in spec_oracle._ranked(t)]
            off.append(len(tax))
        reads = []
        for i, (s, _t) in enumerate(ents):
            for _ in range(12):
                a = rng.randint(0, len(s) - 80)
                reads.append(s[a:a + rng.randint(40, 80)])
        reads += [common.rand_seq(rng, 60) for _ in range(10)]       # UNKNOWN
        data, offsets = g.pack_reads(reads)
        with pytest.raises(Exception):
            g.classify_packed_tax(data, offsets)                     # lineages not set yet
        with pytest.raises(Exception):
            g.set_lineages(off[:-1], tax, flags[:-1])                # must cover every document
        g.set_lineages(off, tax, flags)
        out = g.classify_packed_tax(data, offsets)
        seen = set()
        for i, r in enumerate(reads):
            assert out["status"][i] == 0
            nh = int(out["n_hits"][i])
            entries = [{"taxon_hierarchy": doc_tax[int(d)]} for d in out["top_doc"][i, :nh]]
            typ, rank, name = spec_oracle.label(entries)
            assert out["tax_label"][i] == TYPE[typ], (i, typ, out["tax_label"][i], out["top_doc"][i, :nh])
            ti = int(out["tax_index"][i])
            if rank == "unknown" and name == "":
                assert ti == -1, (i, ti)
            else:
                ranked = spec_oracle._ranked(entries[0]["taxon_hierarchy"])
                assert ti >= 0 and (ranked[ti]["rank"], ranked[ti]["name"]) == (rank, name), (i, ti, rank, name)
            seen.add((typ, rank if rank in ("unknown", "species", "genus") else "other", min(nh, 2)))
        # every branch of makeClassificationResult was exercised
        for need in [("UNKNOWN", "unknown", 0), ("CLASSIFIED", "species", 1), ("CLASSIFIED", "unknown", 1),
                     ("CLASSIFIED", "species", 2), ("CLASSIFIED", "genus", 2), ("VAGUE", "unknown", 2)]:
            assert need in seen, (need, sorted(seen))
    finally:
        g.close()


def test_host_lclassify_device_and_host_lca_write_identical_files(tmp_path, monkeypatch):
    """bsp::LocalClassifier formats from the device labels by default and from its own lineage walk with BSP_HOST_LCA=1:
    the .result files must be byte-identical (malformed taxonomy included: the read is skipped, as the reference's
    exception would)."""
    from biospectra_b200 import _lib
    L = _lib.lib()
    rng = random.Random(23)
    core = [common.rand_seq(rng, 500) for _ in range(3)]
    ents = [(core[0], _lineage(0, 0, 0)), (core[0], _lineage(1, 0, 0)), (core[1], _lineage(2, 1, 0)), (core[1], _lineage(3, 2, 1)),
            (core[2], _lineage(4, 3, 2)), (core[2], "{not json"), (common.rand_seq(rng, 500), ""), (common.rand_seq(rng, 500), _lineage(5, 4, 3))]
    ref = tmp_path / "ref"
    ref.mkdir()
    for i, (s, t) in enumerate(ents):
        (ref / ("%02d.fa" % i)).write_text(">e%d\n%s\n" % (i, s))
        if t:
            (ref / ("%02d.fa.taxd" % i)).write_text(t)
    reads = []
    for i, (s, _t) in enumerate(ents):
        for j in range(40):
            a = rng.randint(0, len(s) - 90)
            reads.append(s[a:a + 60 + j % 30])
    qdir = tmp_path / "q"
    qdir.mkdir()
    (qdir / "r.fa").write_text("".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    idx = tmp_path / "idx"
    conf = tmp_path / "conf.json"
    conf.write_text(json.dumps({"index_path": str(idx), "kmer_size": 8, "kmer_skips": 0, "min_strand_kmer": False,
                                "query_term_min_should_match": 0.5, "worker_threads": 4, "index_ram_buffer": 16,
                                "query_algorithm": "PAIRED_PROXIMITY", "scoring_algorithm": "tfidf"}))
    assert L.bsp_host_index(str(conf).encode(), None, str(ref).encode(), 0) == 0, L.bsp_host_last_error()
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("BSP_HOST_LCA", mode)
        out = tmp_path / ("out" + mode)
        assert L.bsp_host_classify_local(str(conf).encode(), None, str(qdir).encode(), str(out).encode(), 0) == 0, L.bsp_host_last_error()
        outs[mode] = (out / "r.fa.result").read_bytes()
        summ = json.loads((out / "r.fa.result.sum").read_text())
        outs[mode + "s"] = {k: summ[k] for k in ("total", "unknown", "vague", "classified")}
    assert outs["0"] == outs["1"] and outs["0s"] == outs["1s"]
    lines = [json.loads(x) for x in outs["0"].decode().splitlines()]
    types = {x["type"] for x in lines}
    assert {"CLASSIFIED", "VAGUE"} <= types
    assert len(lines) < len(reads)             # reads whose hits include the malformed taxonomy are skipped
