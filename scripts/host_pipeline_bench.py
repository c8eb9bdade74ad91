#!/usr/bin/env python
"""File-to-file throughput of the reference-facing drivers (BioSpectra `index` + `lclassify`, restated in
biospectra_b200/csrc/host over the C ABI): FASTA + .taxd in, <name>.result / <name>.result.sum out.
Small reference on purpose (the GPU part is then negligible): this measures FASTA ingest, batching, label/LCA logic
and the JSON writer -- SURVEY.md 8(f) row 1."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from biospectra_b200 import synth  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_genomes, glen = 10, 2_000_000
    work = tempfile.mkdtemp(prefix="bsp_host_")
    ref, reads_dir, out, idx = (os.path.join(work, x) for x in ("ref", "reads", "out", "idx"))
    for d in (ref, reads_dir):
        os.makedirs(d)
    genomes = [synth.genome_numpy(1, g, glen) for g in range(n_genomes)]
    for g, seq in enumerate(genomes):
        with open(os.path.join(ref, "%02d.fna" % g), "wb") as f:
            f.write(b">syn|%d| synthetic genome %d\n" % (g, g))
            s = seq.tobytes()
            f.write(b"\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + b"\n")
        lineage = [{"taxid": 1000 + g, "name": "strain %d" % g, "parent": 100 + g // 2, "rank": "no rank"},
                   {"taxid": 100 + g // 2, "name": "species %d" % (g // 2), "parent": 10 + g // 4, "rank": "species"},
                   {"taxid": 10 + g // 4, "name": "genus %d" % (g // 4), "parent": 1, "rank": "genus"}]
        with open(os.path.join(ref, "%02d.fna.taxd" % g), "w") as f:
            json.dump(lineage, f)
    reads, src = synth.sample_reads_numpy(genomes, n_reads, 100, 0.04, 11)
    with open(os.path.join(reads_dir, "sample.fa"), "wb") as f:
        f.write(b"".join(b">syn|%d| read %d\n%s\n" % (int(src[i]), i, r) for i, r in enumerate(reads)))
    conf = os.path.join(work, "conf.json")
    with open(conf, "w") as f:
        json.dump({"index_path": idx, "kmer_size": 10, "kmer_skips": 0, "min_strand_kmer": False, "query_term_min_should_match": 0.5,
                   "worker_threads": os.cpu_count() or 4, "index_ram_buffer": 16, "query_algorithm": "PAIRED_PROXIMITY",
                   "scoring_algorithm": "tfidf"}, f)
    cli = os.path.join(ROOT, "biospectra_b200", "biospectra_gpu")
    t0 = time.time()
    subprocess.check_call([cli, "index", "-j", conf, "-r", ref], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t_index = time.time() - t0
    t0 = time.time()
    subprocess.check_call([cli, "lclassify", "-j", conf, "-in", reads_dir, "-out", out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t_cls = time.time() - t0
    summ = json.load(open(os.path.join(out, "sample.fa.result.sum")))
    first = open(os.path.join(out, "sample.fa.result")).readline()
    # the reference's control flow: T pool threads, one Classifier.classify call per read (micro-batched on the GPU side)
    per_read = {}
    for threads in [int(x) for x in os.environ.get("PER_READ_THREADS", "16,256").split(",") if x]:
        out_t = out + "_pr%d" % threads
        subprocess.check_call([cli, "lclassify", "-j", conf, "-in", reads_dir, "-out", out_t, "--per-read", str(threads)],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        st = json.load(open(os.path.join(out_t, "sample.fa.result.sum")))
        assert {k: st[k] for k in ("total", "unknown", "vague", "classified")} == {k: summ[k] for k in ("total", "unknown", "vague", "classified")}
        per_read[str(threads)] = n_reads / ((st["end_time"] - st["start_time"]) / 1e3)
    print(json.dumps({"reads": n_reads, "index_seconds": t_index, "index_mbp_per_s": n_genomes * glen / 1e6 / t_index,
                      "lclassify_seconds_process": t_cls, "lclassify_ms_in_summary": summ["end_time"] - summ["start_time"],
                      "reads_per_s_file_to_file": n_reads / ((summ["end_time"] - summ["start_time"]) / 1e3),
                      "reads_per_s_per_read_calls_by_threads": per_read, "summary": summ, "result_bytes": os.path.getsize(os.path.join(out, "sample.fa.result")),
                      "first_line": first[:300]}))


if __name__ == "__main__":
    main()
