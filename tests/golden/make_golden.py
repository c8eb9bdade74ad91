#!/usr/bin/env python
"""Generates tests/golden/golden_small.json.

The reference (Java + Lucene 5.3.1 jars) cannot run in this image (no JVM) and ships no golden outputs, so these
vectors are produced by the literal spec oracle (oracle/spec_oracle.py), which restates the reference's hot path and
is itself pinned by the hand-traced known answers of SURVEY.md Appendix B (tests/test_oracle_kat.py).  They pin
regressions: the C oracle and the CUDA path (through the C ABI and through the host-side lclassify driver) must
reproduce them bit for bit.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common            # noqa: E402
import spec_oracle as so  # noqa: E402


def bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def lineage(rng, genus, species, strain):
    return json.dumps([{"taxid": 1000 + strain, "name": "strain %d" % strain, "parent": 100 + species, "rank": "no rank"},
                       {"taxid": 100 + species, "name": "species %d" % species, "parent": 10 + genus, "rank": "species"},
                       {"taxid": 10 + genus, "name": "genus %d" % genus, "parent": 1, "rank": "genus"},
                       {"taxid": 1, "name": "root", "parent": 1, "rank": "no rank"}])


CASES = [
    dict(seed=1, k=3, min_strand=False, skips=0, alg="NAIVE_KMER", sim="tfidf", msm=0.5, junk=0.0),
    dict(seed=2, k=5, min_strand=False, skips=0, alg="PAIRED_PROXIMITY", sim="tfidf", msm=0.5, junk=0.02),
    dict(seed=3, k=4, min_strand=True, skips=0, alg="CHAIN_PROXIMITY", sim="tfidf", msm=0.5, junk=0.0),
    dict(seed=4, k=6, min_strand=False, skips=2, alg="PAIRED_PROXIMITY", sim="bm25", msm=0.3, junk=0.01),
    dict(seed=5, k=10, min_strand=False, skips=0, alg="PAIRED_PROXIMITY", sim="tfidf", msm=0.5, junk=0.0),
    dict(seed=6, k=10, min_strand=False, skips=5, alg="PAIRED_PROXIMITY", sim="default", msm=0.5, junk=0.0),
    dict(seed=7, k=8, min_strand=True, skips=0, alg="NAIVE_KMER", sim="bm25", msm=0.9, junk=0.03),
    dict(seed=8, k=12, min_strand=False, skips=1, alg="CHAIN_PROXIMITY", sim="tfidf", msm=0.0, junk=0.0),
]


def main():
    out = []
    for case in CASES:
        rng = random.Random(case["seed"])
        seqs = common.make_corpus(500 + case["seed"], n_entries=(4, 7), length=(20, 260), junk=case["junk"])
        # what FASTAReader#readNext would hand over (upper-cased, trimmed), so that a FASTA round trip is the identity
        seqs = [so.java_trim(x.upper()) for x in seqs]
        seqs = [x if x else "ACGT" for x in seqs]
        entries = []
        for i, s in enumerate(seqs):
            tax = "" if i % 4 == 3 else lineage(rng, i // 4, i // 2, i)
            entries.append({"filename": "ref%d.fa" % (i // 2), "header": "entry|%d| synthetic" % i, "taxonomy": tax, "sequence": s})
        reads = common.make_reads(700 + case["seed"], seqs, n=14, max_len=110)
        ix = so.Index(case["k"], case["min_strand"])
        for e in entries:
            ix.add_entry(e["filename"], e["header"], e["taxonomy"], e["sequence"])
        alg = so.ALGORITHMS[case["alg"]]
        sim = so.BM25 if case["sim"].lower() == "bm25" else so.TFIDF
        results = []
        for qi, q in enumerate(reads):
            r = so.search(ix, q, skips=case["skips"], algorithm=alg, scoring=sim, msm_ratio=case["msm"])
            if r is None:
                results.append(None)
                continue
            hits, nc, msm = r
            cj = so.classify(ix, ">read%d" % qi, q, skips=case["skips"], algorithm=alg, scoring=sim, msm_ratio=case["msm"])
            results.append({"n_clauses": nc, "msm": msm, "top": [[d, bits(s)] for d, s in hits], "result": cj})
        terms = sorted(ix.postings)
        sample = terms[:: max(1, len(terms) // 12)][:12]
        out.append({
            "config": case, "entries": entries, "reads": reads,
            "docs": [{"direction": d.direction, "entry_header": d.header, "ntok": d.ntok, "norm": d.norm} for d in ix.docs],
            "n_terms": len(ix.postings), "n_postings": sum(len(v) for v in ix.postings.values()),
            "n_positions": sum(d.ntok for d in ix.docs),
            "term_samples": [{"key": t, "postings": [[d, ix.postings[t][d]] for d in sorted(ix.postings[t])]} for t in sample],
            "results": results,
        })
    path = os.path.join(HERE, "golden_small.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/spec_oracle.py; the reference cannot run here)",
                   "cases": out}, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
