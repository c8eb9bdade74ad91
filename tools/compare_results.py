#!/usr/bin/env python
"""Compare two BioSpectra `.result` files (one JSON object per line, beans/ClassificationResult.java:59-117): the
reference's (Lucene) and this library's.  Documents are matched by (filename, header, sequence_direction) -- Lucene's doc
ids depend on its indexing threads, this library's are canonical -- and reads by query_header.

north_star tiers: scores within 1e-5 relative; type / taxon and the hit set equal on >= 99.99 % of the reads, every
disagreement listed (and expected to be a near-threshold tie).  Exit status 0 iff both hold.

    python tools/compare_results.py <reference.result> <gpu.result> [--rel 1e-5] [--agree 0.9999]
"""
from __future__ import annotations

import argparse
import json
import sys


def load(path):
    out = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                d = json.loads(line)
                out[d["query_header"]] = d
    return out


def hit_key(e):
    return (e.get("filename"), e.get("header"), e.get("sequence_direction"))


def compare(ref: dict, got: dict, rel: float = 1e-5):
    """-> dict(reads, missing, extra, label_mismatch, hitset_mismatch, score_mismatch, max_rel, disagreements[list])"""
    rep = {"reads": len(ref), "missing": [], "extra": [], "disagreements": [], "score_mismatch": 0, "max_rel": 0.0}
    for h in got:
        if h not in ref:
            rep["extra"].append(h)
    for h, r in ref.items():
        g = got.get(h)
        if g is None:
            rep["missing"].append(h)
            continue
        why = []
        if (r.get("type"), r.get("taxon_rank"), r.get("taxon_name")) != (g.get("type"), g.get("taxon_rank"), g.get("taxon_name")):
            why.append("label %r vs %r" % ((r.get("type"), r.get("taxon_rank"), r.get("taxon_name")),
                                           (g.get("type"), g.get("taxon_rank"), g.get("taxon_name"))))
        rs = {hit_key(e): e for e in (r.get("result") or [])}
        gs = {hit_key(e): e for e in (g.get("result") or [])}
        if set(rs) != set(gs):
            why.append("hit set differs: only in reference %s, only here %s" % (sorted(set(rs) - set(gs))[:3], sorted(set(gs) - set(rs))[:3]))
        worst = 0.0
        for k in set(rs) & set(gs):
            a, b = float(rs[k]["score"]), float(gs[k]["score"])
            d = abs(a - b) / max(abs(a), abs(b), 1e-30)
            worst = max(worst, d)
        # scores of a read's hits are all equal to its maximum, so the hits missing on one side tell how close the tie was
        rep["max_rel"] = max(rep["max_rel"], worst)
        if worst > rel:
            rep["score_mismatch"] += 1
            why.append("score differs by %.3g relative" % worst)
        if why:
            top_r = max([float(e["score"]) for e in rs.values()] or [0.0])
            top_g = max([float(e["score"]) for e in gs.values()] or [0.0])
            rep["disagreements"].append({"query_header": h, "why": why, "top_score_reference": top_r, "top_score_here": top_g})
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("ours")
    ap.add_argument("--rel", type=float, default=1e-5)
    ap.add_argument("--agree", type=float, default=0.9999)
    a = ap.parse_args()
    rep = compare(load(a.reference), load(a.ours), a.rel)
    n = max(rep["reads"], 1)
    bad = len(rep["disagreements"]) + len(rep["missing"])
    agree = 1.0 - bad / n
    print(json.dumps({"reads": rep["reads"], "missing_here": len(rep["missing"]), "extra_here": len(rep["extra"]),
                      "disagreements": len(rep["disagreements"]), "score_mismatch": rep["score_mismatch"],
                      "max_relative_score_difference": rep["max_rel"], "agreement": agree}))
    for d in rep["disagreements"][:200]:
        print(json.dumps(d))
    for h in rep["missing"][:50]:
        print(json.dumps({"query_header": h, "why": ["no line here (read dropped?)"]}))
    ok = agree >= a.agree and rep["score_mismatch"] == 0 and not rep["extra"]
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
