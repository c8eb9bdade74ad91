#!/usr/bin/env python
"""Writes tests/golden/reference_reads.json.gz: a deterministic subset of the ONLY test data the reference ships
 -- its query fixtures:

  /root/reference/benchmark/bacteria_ncbi_simulated/sample_0.0.fa ... sample_0.19.fa
        20 x 10,000 simulated reads (90-109 bp, 0-19 % substitutions); header = the source genome's FASTA header
  /root/reference/benchmark/simHC.20.500/*.fna
        20 x 500 Sanger reads of the FAMeS simHC set (155-1,791 bp) with X-masked vector and N calls

The reference genomes these reads were drawn from are not in the repository (NCBI download), so the tests build a
synthetic index out of the reads themselves (tests/test_reference_reads.py) and compare the CUDA path with the oracle on
it.  /root/reference does not exist on the GPU box, hence this committed subset; run this script (in the build
container) to regenerate it.  Entries are parsed with the reference reader's semantics (oracle/spec_oracle.read_fasta =
org.yeastrc.fasta.FASTAReader#readNext, executed from the jar in tests/golden/fasta_reader.json)."""
import gzip
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spec_oracle as so  # noqa: E402

REF = "/root/reference/benchmark"
PER_SAMPLE = 40       # reads per sample_*.fa
PER_SIMHC = 16        # reads per simHC file


def main():
    rng = random.Random(20)
    out = {"simulated": [], "simhc": [], "source": "iychoi/biospectra benchmark/bacteria_ncbi_simulated, benchmark/simHC.20.500"}
    sim_dir = os.path.join(REF, "bacteria_ncbi_simulated")
    for fn in sorted(os.listdir(sim_dir), key=lambda x: float(x[len("sample_"):-3])):
        ents = so.read_fasta(open(os.path.join(sim_dir, fn), encoding="latin-1").read())
        for i in sorted(rng.sample(range(len(ents)), PER_SAMPLE)):
            h, s = ents[i]
            out["simulated"].append({"file": fn, "index": i, "header": h[1:], "seq": s})
    hc_dir = os.path.join(REF, "simHC.20.500")
    for fn in sorted(os.listdir(hc_dir)):
        ents = so.read_fasta(open(os.path.join(hc_dir, fn), encoding="latin-1").read())
        order = sorted(range(len(ents)), key=lambda i: -len(ents[i][1]))
        pick = set(order[:4]) | set(order[-2:])                                   # longest (> 510 clauses) and shortest
        masked = [i for i in range(len(ents)) if "X" in ents[i][1] and "N" in ents[i][1]]
        pick |= set(masked[:4])
        rest = [i for i in range(len(ents)) if i not in pick]
        pick |= set(rng.sample(rest, PER_SIMHC - len(pick)))
        for i in sorted(pick):
            h, s = ents[i]
            out["simhc"].append({"file": fn, "index": i, "header": h[1:], "seq": s})
    path = os.path.join(ROOT, "tests", "golden", "reference_reads.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print(path, os.path.getsize(path), "bytes;", len(out["simulated"]), "simulated reads,", len(out["simhc"]), "simHC reads, longest",
          max(len(r["seq"]) for r in out["simhc"]))


if __name__ == "__main__":
    main()
