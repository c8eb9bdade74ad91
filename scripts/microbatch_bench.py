"""Single-read entry point under concurrency: T threads call Classifier.classify_one in a loop,
with coalescing on (default window) and off (max_reads = 1).  Run on a GPU box:
    python scripts/microbatch_bench.py [--threads 64] [--reads 20000]
Python threads release the GIL inside the C call; the interpreter still serialises the call overhead (~10 us), so the
numbers are a floor for what a JNI caller sees."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=64)
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--genome-bp", type=int, default=1_000_000)
    args = ap.parse_args()
    from biospectra_b200 import Configuration, Indexer
    rng = np.random.default_rng(5)
    ix = Indexer(Configuration(index_path="", kmer_size=10))
    genomes = []
    for g in range(args.genomes):
        s = rng.integers(0, 4, args.genome_bp, dtype=np.uint8)
        s = np.frombuffer(b"ACGT", np.uint8)[s].tobytes()
        genomes.append(s)
        ix.add_entry("%d.fna" % g, "syn|%d|" % g, "", s.decode())
    clf = ix.close_to_classifier(Configuration(index_path="", kmer_size=10, kmer_skips=0, query_algorithm="PAIRED_PROXIMITY"))
    reads = []
    for i in range(args.reads):
        g = genomes[int(rng.integers(0, len(genomes)))]
        a = int(rng.integers(0, len(g) - 100))
        reads.append(g[a:a + 100])
    out = {}
    for name, (mx, us) in {"coalesced_default": (4096, 200), "coalesced_no_wait": (4096, 0), "one_read_per_batch": (1, 0)}.items():
        clf.set_microbatch(mx, us)
        b0 = clf.microbatch_stats()
        n = len(reads) if mx > 1 else min(len(reads), 2000)

        def work(t):
            for i in range(t, n, args.threads):
                clf.classify_one(reads[i])

        th = [threading.Thread(target=work, args=(t,)) for t in range(args.threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        b1 = clf.microbatch_stats()
        out[name] = {"reads_per_s": n / dt, "reads": n, "gpu_batches": b1["batches"] - b0["batches"],
                     "mean_batch": n / max(1, b1["batches"] - b0["batches"])}
    # latency of a lone caller
    clf.set_microbatch(4096, 0)
    t0 = time.perf_counter()
    for i in range(200):
        clf.classify_one(reads[i])
    out["lone_caller_latency_us"] = (time.perf_counter() - t0) / 200 * 1e6
    out["threads"] = args.threads
    print(json.dumps(out))
    clf.close()


if __name__ == "__main__":
    main()
