#!/usr/bin/env python
"""Small synthetic corpus for the side-by-side run against the real reference (tools/ref_lucene_run.sh): FASTA files
with `.taxd` sidecars (format of the reference's tools/gen_taxon_hierarchy.py), reads drawn like the reference's
simulator (verification/RandomReadSampler.java:99-149), and two copies of the shipped configuration.json that differ
only in index_path.  Deterministic.

    python tools/make_parity_corpus.py <workdir> [--genomes 12] [--genome-bp 200000] [--reads 5000] [--skips 0]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workdir")
    ap.add_argument("--genomes", type=int, default=12)
    ap.add_argument("--genome-bp", type=int, default=200_000)
    ap.add_argument("--reads", type=int, default=5000)
    ap.add_argument("--skips", type=int, default=0)
    ap.add_argument("--algorithm", default="PAIRED_PROXIMITY")
    a = ap.parse_args()
    from biospectra_b200 import synth
    w = a.workdir
    ref, reads_dir = os.path.join(w, "ref"), os.path.join(w, "reads")
    for d in (ref, reads_dir):
        os.makedirs(d, exist_ok=True)
    genomes = [synth.genome_numpy(1, g, a.genome_bp) for g in range(a.genomes)]
    # related strains, 99.9 % identical (most 100 bp windows are shared, so their reads tie exactly): every third genome
    # copies its predecessor (same species: the tie classifies at the species through the common-taxon rule), and the
    # last genome copies genome 0 (another genus once there are more than 6 genomes: nothing ranked in common -> VAGUE)
    import numpy as np
    rng = np.random.default_rng(7)

    def near_copy(src):
        s = src.copy()
        pos = rng.choice(len(s), max(1, len(s) // 1000), replace=False)
        s[pos] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(pos))]
        return s
    for g in range(2, a.genomes - 1, 3):
        genomes[g] = near_copy(genomes[g - 1])
    if a.genomes > 1:
        genomes[a.genomes - 1] = near_copy(genomes[0])
    for g, seq in enumerate(genomes):
        with open(os.path.join(ref, "%02d.fna" % g), "wb") as f:
            f.write(b">syn|%d| synthetic genome %d\n" % (g, g))
            s = seq.tobytes()
            f.write(b"\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + b"\n")
        lineage = [{"taxid": 1000 + g, "name": "strain %d" % g, "parent": 100 + g // 3, "rank": "no rank"},
                   {"taxid": 100 + g // 3, "name": "species %d" % (g // 3), "parent": 10 + g // 6, "rank": "species"},
                   {"taxid": 10 + g // 6, "name": "genus %d" % (g // 6), "parent": 1, "rank": "genus"}]
        with open(os.path.join(ref, "%02d.fna.taxd" % g), "w") as f:
            json.dump(lineage, f)
    reads, src = synth.sample_reads_numpy(genomes, a.reads, 100, 0.04, 11)
    with open(os.path.join(reads_dir, "sample.fa"), "wb") as f:
        f.write(b"".join(b">read%d|syn|%d|\n%s\n" % (i, int(src[i]), r) for i, r in enumerate(reads)))
    for tag in ("ref", "gpu"):
        with open(os.path.join(w, "conf_%s.json" % tag), "w") as f:
            json.dump({"index_path": os.path.join(w, "index_" + tag), "kmer_size": 10, "kmer_skips": a.skips, "min_strand_kmer": False,
                       "query_term_min_should_match": 0.5, "worker_threads": 4, "index_ram_buffer": 16,
                       "query_algorithm": a.algorithm, "scoring_algorithm": "tfidf"}, f)
    print(json.dumps({"workdir": w, "genomes": a.genomes, "reads": a.reads}))


if __name__ == "__main__":
    main()
