#!/usr/bin/env python3
"""Per-CUDA-line totals of an `ncu --page source --csv --print-source cuda,sass` export: instructions executed and warp
samples, by file and line.  Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | tools/ncu_lines.py [min_pct]"""
import csv
import sys

def main():
    min_pct = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
    rows = list(csv.reader(sys.stdin))
    cur_file, hdr, acc, src = None, None, {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or r[0] in ("Function Name",):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        key = (cur_file, ln)
        src[key] = r[1]
        try:
            n, s = int(r[ii] or 0), int(r[si] or 0)
        except (ValueError, IndexError):
            continue
        a = acc.setdefault(key, [0, 0])
        a[0] += n
        a[1] += s
    tot_i = sum(a[0] for a in acc.values()) or 1
    tot_s = sum(a[1] for a in acc.values()) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key in sorted(acc):
        n, s = acc[key]
        if 100.0 * n / tot_i >= min_pct or 100.0 * s / tot_s >= min_pct:
            print("%-14s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (key[0], key[1], 100.0 * n / tot_i, 100.0 * s / tot_s, src[key].strip()[:100]))

if __name__ == "__main__":
    main()
