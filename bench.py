#!/usr/bin/env python
"""bench.py -- reads/s classified on the BASELINE.json workload.

Default (N=1): config C2 of BASELINE.json -- ~3,000 synthetic genomes / ~10 Gbp indexed on one B200,
1 M reads of ~100 bp at 4 % error, configuration.json as shipped by the reference
(k=10, skips=0, msm=0.5, PAIRED_PROXIMITY, tfidf).  One "step" = one pass of the classify hot path over
the whole read batch.  N>1 (torchrun): read-sharded replicas, weak scaling, no collective on the data path.

  value : reads/s with reads already resident in HBM (bsp_classify_device)
  e2e   : reads/s through the reference-facing call with HOST buffers (bsp_classify_packed: H2D of the
          reads + D2H of labels/top-10 inside the timed region)
  --impl reference : the reference's CPU implementation of the path.  No JVM exists in this image, so
          this is the C/OpenMP oracle restatement (oracle/bsp_oracle.c) on all host cores, on a bounded
          sample (see cpu_baseline.sample), measured at three index sizes so that the scaling to the full
          document count is a fit, not an assumption.

The same JSON line also carries (N=1) `secondary`: the same index queried with the reference's other settings
(CHAIN_PROXIMITY, NAIVE_KMER, the code-default kmer_skips=5, bm25) and a k=16 index, and `build`: the index
build with its byte model; at N>1 `docpart`: the same reference split N ways (doc-partitioned mode, one NCCL
all_gather per batch), verified against the replica on a read subset.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: genomes, total bp, reads per GPU, read length, error
    "C2": dict(genomes=3000, total_bp=10_000_000_000, reads=1_000_000, R=100, err=0.04, seed=2, read_seed=12),
    "C1": dict(genomes=10, total_bp=20_000_000, reads=10_000, R=100, err=0.04, seed=1, read_seed=11, equal=True),
    "mid": dict(genomes=300, total_bp=1_000_000_000, reads=200_000, R=100, err=0.04, seed=4, read_seed=14),
    # BASELINE configs[3]: larger than one GPU; only with --mode docpart on 8 GPUs (7.5 Gbp per GPU)
    "C4": dict(genomes=18000, total_bp=60_000_000_000, reads=1_000_000, R=100, err=0.04, seed=3, read_seed=13),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def measured_traffic(args, kernel):
    """dram__bytes_read+write per launch of the dominant kernel from the committed `ncu --set full` capture of the same kernel
    on the same workload (profiles/r2_ncu_kernel_metrics.json, else round 1's); only valid for that workload"""
    if args.workload != "C2" or kernel != "filter_score_kernel" or args.algorithm != "PAIRED_PROXIMITY":
        return None
    for fn, key in (("r2_ncu_kernel_metrics.json", "r2_filter_final_C2"), ("r1_ncu_kernel_metrics.json", "r1j_filter_final_C2")):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                m = json.load(f)[key]

            def gb(k):
                v, u = float(m[k]["value"]), m[k]["unit"]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            return gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        except Exception:
            continue
    return None


def shipped_config(**kw):
    from biospectra_b200 import Configuration
    # /root/reference/configuration.json as shipped
    c = Configuration(index_path="", kmer_size=10, kmer_skips=0, min_strand_kmer=False, query_term_min_should_match=0.5,
                      worker_threads=4, index_ram_buffer=16, query_algorithm="PAIRED_PROXIMITY", scoring_algorithm="tfidf")
    for k, v in kw.items():
        setattr(c, k, v)
    return c


ALG_ID = {"NAIVE_KMER": 0, "CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}


def sample_plan(wl, lengths):
    """CPU baseline samples: the first S, S/2 and S/4 genomes (S: <= ~120 Mbp) of the same reference, reads drawn from them."""
    budget_bp = int(os.environ.get("BSP_CPU_SAMPLE_BP", 120_000_000))
    S = 1
    while S < len(lengths) and lengths[:S + 1].sum() <= budget_bp:
        S += 1
    return S


def cpu_threads():
    """threads of the CPU arm: every host core unless BSP_CPU_THREADS pins it (reported as cpu_baseline.cores)"""
    return int(os.environ.get("BSP_CPU_THREADS", 0)) or (os.cpu_count() or 1)


class CpuArm:
    """The reference's CPU implementation of the path, as far as it can exist here: the Java/Lucene reference cannot
    run (no JVM in the image), so this is the C/OpenMP oracle restatement (oracle/bsp_oracle.c, pinned to the reference's
    own Lucene bytecode by tests/golden/lucene_*.json) on the host cores.  Bounded sample: index = the first S genomes of
    the same reference (<= ~120 Mbp: the oracle's index build is sort-based and slow), reads drawn from them with the
    same simulator."""

    def __init__(self, wl, lengths, kept, S, n_reads, threads=None, algorithm=2, skips=0, scoring=0, min_strand=False, k=10):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_lib as ol
        from biospectra_b200 import synth
        self.threads = threads or cpu_threads()
        self.q = dict(skips=skips, algorithm=algorithm, scoring=scoring, msm=0.5)
        self.S, self.G = S, len(lengths)
        t0 = time.time()
        self.ox = ol.OracleIndex(k, bool(min_strand))
        gl = []
        for g in sorted(kept)[:S]:
            self.ox.add_entry(kept[g])
            gl.append(np.frombuffer(kept[g], np.uint8))
        self.ox.finish()
        self.index_seconds = time.time() - t0
        self.n_docs = int(self.ox.stats()["n_docs"])
        self.sample_bp = int(sum(len(x) for x in gl))
        self.reads, self.src = synth.sample_reads_numpy(gl, n_reads, wl["R"], wl["err"], wl["read_seed"] + 1000)
        self.classify(self.reads[:256])                                                                   # warm

    def classify(self, reads, **over):
        return self.ox.classify(reads, threads=self.threads, **dict(self.q, **over))

    def step(self):
        t0 = time.time()
        res = self.classify(self.reads)
        return time.time() - t0, res


def fit_linear(docs, us_per_read):
    """least squares us_per_read = a + b * docs; returns (a, b, r2)"""
    x, y = np.asarray(docs, np.float64), np.asarray(us_per_read, np.float64)
    if len(x) < 2 or np.ptp(x) == 0:
        return 0.0, float(y[0] / x[0]), None
    b, a = np.polyfit(x, y, 1)
    pred = a + b * x
    ss_res, ss_tot = float(((y - pred) ** 2).sum()), float(((y - y.mean()) ** 2).sum())
    return float(a), float(b), (1.0 - ss_res / ss_tot if ss_tot > 0 else None)


def cpu_arm_points(wl, lengths, kept, S, n_reads, algorithm, skips=0, scoring=0, largest=None):
    """per-read time of the CPU arm at three index sizes (S/4, S/2, S genomes) -> points + the arm of the largest"""
    sizes = sorted({max(1, S // 4), max(1, S // 2), S})
    points, arm = [], None
    for sz in sizes:
        a = largest if (largest is not None and sz == S) else CpuArm(wl, lengths, kept, sz, max(2000, n_reads * sz // S), algorithm=algorithm,
                                                                     skips=skips, scoring=scoring)
        dt, res = a.step()
        points.append({"genomes": sz, "docs": a.n_docs, "reads": len(a.reads), "reads_per_s": len(a.reads) / dt,
                       "us_per_read": dt / len(a.reads) * 1e6, "index_seconds": a.index_seconds})
        if sz == S:
            arm, last = a, (dt, res)
    return points, arm, last


def cpu_report(points, arm, dt, res, wl_docs):
    """cpu_baseline object: the measured points, the linear fit of per-read time against the document count, and the
    workload-size value the fit gives (per-read cost at k=10 is a walk over every document's postings: linear in docs)"""
    a, b, r2 = fit_linear([p["docs"] for p in points], [p["us_per_read"] for p in points])
    us_full = a + b * wl_docs
    n = len(arm.reads)
    return {"value": 1e6 / us_full, "unit": "reads/s", "cores": arm.threads, "kind": "port",
            "sample": "C/OpenMP oracle restatement of the reference's Lucene path (no JVM in this image; pinned to the reference's Lucene "
                      "5.3.1 bytecode by tests/golden/lucene_*.json), %d threads: indexes of the first %s of %d genomes of the same reference "
                      "(largest: %.0f Mbp, %d docs, %d reads drawn from it: %.1f reads/s measured); value = 1 / (a + b * %d docs) from the "
                      "least-squares fit of per-read time against the document count over the three sizes"
                      % (arm.threads, "/".join(str(p["genomes"]) for p in points), arm.G, arm.sample_bp / 1e6, arm.n_docs, n, n / dt, wl_docs),
            "measured_points": points,
            "fit": {"a_us_per_read": a, "b_us_per_read_per_doc": b, "r2": r2, "workload_docs": wl_docs, "us_per_read_at_workload": us_full,
                    "proportional_scaling_value": (n / dt) * arm.n_docs / float(wl_docs)},
            "measured_reads_per_s_on_sample": n / dt, "index_seconds": arm.index_seconds, "classify_seconds": dt,
            "classified_fraction": float((res["label"] == 2).mean())}


def compare_results(res_g, res_o):
    same = (res_g["status"] == res_o["status"]) & (res_g["label"] == res_o["label"]) & (res_g["n_hits"] == res_o["n_hits"])
    for j in range(10):
        live = res_o["n_hits"] > j
        same &= ~live | ((res_g["top_doc"][:, j] == res_o["top_doc"][:, j]) &
                         (res_g["top_score"][:, j].view(np.uint32) == res_o["top_score"][:, j].view(np.uint32)))
    return same


def gpu_parity_on_sample(arm, res_o, kept, wl, lengths, device, main_kw):
    """The oracle is the checker here, never the thing measured: index the CPU arm's sample genomes with the CUDA library
    and require bit-identical results (status, label, hit count, hit doc ids and score bit patterns) on the arm's reads --
    for the benchmarked configuration and for the reference's other settings (code-default kmer_skips=5, bm25, NAIVE_KMER,
    CHAIN_PROXIMITY), plus min_strand_kmer=true on the smallest sample (it is an index-side setting)."""
    from biospectra_b200 import Indexer
    out = {}

    def gpu_index(S, min_strand=False):
        ix = Indexer(shipped_config(device=device, min_strand_kmer=min_strand))
        for g in sorted(kept)[:S]:
            ix.add_entry("%d.fna" % g, "syn|%d|" % g, "", kept[g])
        return ix.close_to_classifier(shipped_config(device=device, min_strand_kmer=min_strand, **main_kw))

    clf = gpu_index(arm.S)
    try:
        m = min(len(arm.reads), 20000)
        cases = [("benchmarked", dict(main_kw), None)]
        for name, kw, okw in (("kmer_skips=5", dict(kmer_skips=5), dict(skips=5)), ("bm25", dict(scoring_algorithm="bm25"), dict(scoring=1)),
                              ("NAIVE_KMER", dict(query_algorithm="NAIVE_KMER"), dict(algorithm=0)),
                              ("CHAIN_PROXIMITY", dict(query_algorithm="CHAIN_PROXIMITY"), dict(algorithm=1))):
            cases.append((name, kw, okw))
        for name, kw, okw in cases:
            clf.set_query(shipped_config(device=device, **kw))
            reads = arm.reads if okw is None else arm.reads[:m]
            ro = res_o if okw is None else arm.classify(reads, **okw)
            for filt in (("1",) if okw is None else ("1", "0")):          # the filter kernel, and the kernel the router picks without it
                os.environ["BSP_FILTER"] = filt
                try:
                    rg = clf.classify_batch(reads)
                finally:
                    os.environ.pop("BSP_FILTER", None)
                same = compare_results(rg, ro)
                key = name if filt == "1" else name + " (BSP_FILTER=0)"
                out[key] = {"reads": len(reads), "identical": int(same.sum()), "mismatching_reads": np.nonzero(~same)[0][:10].tolist()}
            if okw is None:
                # tools/verify_simulator_result.py:10-39 -- the figure bench.py reports at full size, here with the oracle beside it
                src_doc = 2 * np.asarray(arm.src, np.int64)
                out[name]["source_doc_is_top_hit"] = {"gpu": float((rg["top_doc"][:, 0] == src_doc).mean()),
                                                      "oracle": float((ro["top_doc"][:, 0] == src_doc).mean())}
    finally:
        clf.close()
    # min_strand_kmer = true: one document per entry, canonical k-mers (index-side setting: its own small index on both sides)
    S4 = max(1, arm.S // 4)
    ms = CpuArm(wl, lengths, kept, S4, 5000, algorithm=ALG_ID[main_kw.get("query_algorithm", "PAIRED_PROXIMITY")], min_strand=True)
    clf = gpu_index(S4, min_strand=True)
    try:
        ro = ms.classify(ms.reads)
        os.environ["BSP_FILTER"] = "1"
        try:
            rg = clf.classify_batch(ms.reads)
        finally:
            os.environ.pop("BSP_FILTER", None)
        same = compare_results(rg, ro)
        out["min_strand_kmer"] = {"reads": len(ms.reads), "docs": ms.n_docs, "identical": int(same.sum()),
                                  "mismatching_reads": np.nonzero(~same)[0][:10].tolist()}
    finally:
        clf.close()
    return out


def gen_sample_genomes(wl, lengths, S):
    """the first S genomes of the workload, byte-identical to what the GPU arm indexes (same CUDA Philox streams)"""
    import torch
    from biospectra_b200 import synth
    if torch.cuda.is_available():
        r = synth.build_reference_and_reads_cuda(None, lengths[:S], 0, wl["R"], wl["err"], wl["seed"], wl["read_seed"],
                                                 torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))), keep_genomes=set(range(S)))
        return r["kept"]
    return {g: synth.genome_numpy(wl["seed"], g, int(lengths[g])).tobytes() for g in range(S)}


def run_reference(args, wl, lengths):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    S = sample_plan(wl, lengths)
    kept = gen_sample_genomes(wl, lengths, S)
    n_reads = int(os.environ.get("BSP_CPU_SAMPLE_READS", 200000))
    alg = ALG_ID[args.algorithm]
    points, arm, _ = cpu_arm_points(wl, lengths, kept, S, n_reads, alg, skips=args.skips)
    for _ in range(args.warmup):
        arm.step()
    steps = [arm.step() for _ in range(args.steps)]
    dt = float(np.mean([x[0] for x in steps]))
    points[-1].update(reads_per_s=len(arm.reads) / dt, us_per_read=dt / len(arm.reads) * 1e6)
    rep = cpu_report(points, arm, dt, steps[-1][1], 2 * wl["genomes"])
    v = rep["value"]
    line = {"impl": "reference", "metric": "reads/sec classified", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 scores, f64 accumulate", "data": "synthetic",
            "config": workload_config(args, wl), "cpu_baseline": rep,
            "step_note": "one step = the CPU arm over its %d sample reads on the %d-doc sample index (ms_per_step); value is the fit's "
                         "reads/s at the workload's %d docs" % (len(arm.reads), arm.n_docs, 2 * wl["genomes"]),
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, wl):
    shipped = args.k == 10 and args.skips == 0 and args.algorithm == "PAIRED_PROXIMITY" and getattr(args, "scoring", "tfidf") == "tfidf"
    return {"workload": "%s: %d synthetic genomes / %.2f Gbp indexed per GPU (k=%d, both strands), %d reads of ~%d bp at %.0f%% "
                        "error per GPU; %s (skips=%d, msm=0.5, %s, %s)"
                        % (args.workload, wl["genomes"], wl["total_bp"] / 1e9, args.k, wl["reads"], wl["R"], wl["err"] * 100,
                           "configuration.json as shipped" if shipped else "configuration.json with overrides", args.skips, args.algorithm,
                           getattr(args, "scoring", "tfidf")),
            "parallelism": "read-sharded replicas x%d (no collective)" % args.gpus,
            "l2": "inputs larger than L2: the dense index rows read per step span the whole table (>> 126 MB)"}


def run_docpart(args, wl, lengths):
    """Doc-partitioned mode (BASELINE configs[3] style): rank r holds the documents of a contiguous slice of the genomes,
    every rank scores the SAME reads against its slice; build: all_gather of (docs, tokens) + all_reduce of the dense df
    array; query: one all_gather of the per-rank hit lists + merge (biospectra_b200/partition.py).  value = reads/s of
    the whole job against the whole (partitioned) index.  With --verify rank 0 also builds the whole index on its own
    and requires the merged results to be identical to the single-index results."""
    import torch
    import torch.distributed as dist
    from biospectra_b200 import Indexer, partition, synth
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    plan = partition.plan_entries(lengths, world)
    e0, e1 = plan[rank]
    qconf = shipped_config(device=local, query_algorithm=args.algorithm)
    t0 = time.time()
    ix = Indexer(shipped_config(device=local))
    full = Indexer(shipped_config(device=local)) if (args.verify and rank == 0) else None

    class Both:                      # rank 0 with --verify feeds the whole-index handle as well
        def add_entry_device(self, *a):
            g = int(a[0].split(".")[0])
            if e0 <= g < e1:
                ix.add_entry_device(*a)
            if full is not None:
                full.add_entry_device(*a)
    data = synth.build_reference_and_reads_cuda(Both(), lengths, wl["reads"], wl["R"], wl["err"], wl["seed"], wl["read_seed"], dev)
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    clf = ix.close_to_classifier(qconf)
    dp = partition.DocPartitionedClassifier(partition.GpuPartitionEngine(clf, dev)).sync_statistics()
    torch.cuda.synchronize()
    t_build = time.time() - t0
    n = len(data["offsets"]) - 1

    def step():
        return dp.classify(data["seq_bytes"], data["offsets"])

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    torch.cuda.synchronize()
    el = time.perf_counter() - t
    tt = torch.tensor([el], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    el = float(tt[0])
    verify = None
    if full is not None:
        fclf = full.close_to_classifier(qconf)
        ref = fclf.classify_packed(data["seq_bytes"], data["offsets"])
        got = {k: v.cpu().numpy() for k, v in res.items()}
        same = (got["status"] == ref["status"]) & (got["label"] == ref["label"]) & (got["n_hits"] == ref["n_hits"])
        for j in range(10):
            live = ref["n_hits"] > j
            same &= ~live | ((got["top_doc"][:, j] == ref["top_doc"][:, j]) &
                             (got["top_score"][:, j].view(np.uint32) == ref["top_score"][:, j].view(np.uint32)))
        verify = {"reads": n, "identical_to_single_index": int(same.sum())}
        fclf.close()
    if rank == 0:
        print(json.dumps({"metric": "reads/sec classified", "mode": "doc-partitioned", "value": n * args.steps / el, "unit": "reads/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 scores, f64 accumulate",
                          "data": "synthetic",
                          "config": {"workload": "%s: %d genomes / %.2f Gbp split over %d GPUs by contiguous genome ranges, %d reads scored "
                                                 "on every GPU, hit lists merged after one NCCL all_gather" %
                                                 (args.workload, wl["genomes"], wl["total_bp"] / 1e9, world, n),
                                     "parallelism": "doc-partitioned x%d" % world},
                          "partition": {"plan": plan, "global_docs": dp.global_docs, "gen_seconds": t_gen, "build_seconds": t_build},
                          "verify": verify}))
    clf.close()
    dist.barrier()
    dist.destroy_process_group()


def timed_steps(fn, steps, sync, clf=None):
    """wall clock of exactly `steps` calls between synchronisation points (+ the library's CUDA-event kernel/pipeline ms)"""
    sync()
    t = time.perf_counter()
    kernel_ms = pipe_ms = 0.0
    launches = 0
    for _ in range(steps):
        fn()
        if clf is not None:
            tm = clf.last_timings()
            kernel_ms += tm["score_kernel_ms"]
            pipe_ms += tm["pipeline_ms"]
            launches += clf.last_counters()["launches"]
    sync()
    return time.perf_counter() - t, kernel_ms, pipe_ms, launches


def gap_join_traffic(gj):
    """dram read + write bytes of the committed `--set full` capture of gap_join_kernel (profiles/r2_ncu_gap_join.json), valid when
    the launch did the same work as the captured one (same position bytes within 2 %: C2, kmer_skips = 5, 1 M reads)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_gap_join.json")) as f:
            m = json.load(f)["r2gj_gap_join_final_full_1m_reads"]

        def gb(k):
            v, u = float(m[k]["value"]), m[k]["unit"]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        if abs(gj["position_bytes"] - 675542134176) > 0.02 * 675542134176:
            return None
        return gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    except Exception:
        return None


def gap_join_roofline(gj, peak, peak_src):
    """gap_join_kernel (gapjoin.cuh): achieved = position bytes the launch fetched (counted by the kernel: the first term's
    slice once per chunk and segment, every clause's second-term slice once per segment) / its CUDA-event time"""
    s = gj["join_kernel_ms"] / 1e3
    ach = gj["position_bytes"] / s / 1e9 if s > 0 else 0.0
    traffic = gap_join_traffic(gj)
    return {"bound": "hbm", "kernel": "gap_join_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_frac": (traffic / s / 1e9 / peak) if (traffic and s > 0) else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": gj["position_bytes"], "kernel_ms_per_launch": gj["join_kernel_ms"],
            "byte_model": "4 B x positions: per (chunk of <= 8 clauses sharing the first k-mer, segment) that k-mer's slice once + every "
                          "clause's second k-mer slice; directory lookups (8 B per term and segment) not counted",
            "clauses": gj["clauses"], "chunks": gj["chunks"], "matches": gj["matches"], "candidates": gj.get("candidates"),
            "verify_kernel_ms_per_launch": gj.get("verify_kernel_ms")}


def route_roofline(cnt, cost, layout, n_docs, total_read_bytes, kernel_s, peak, peak_src, traffic=None, gap_join=None):
    """Roofline of the scoring kernel that took most reads of the launch.  Algorithmic bytes per launch:
      filter / exact kernels (dense 4-bit rows): rows = clauses x row_bytes (must come from DRAM: 15.8 GB table, random rows),
          io = read bytes + 80 B of results per read; the per-doc weights (4 B x docs per read) are L2-resident and are
          reported separately, NOT counted in `achieved`;
      positional kernels (CSR postings): SURVEY.md 8(d)'s A_read = len + sum over unique terms [16 + 8 df + 4 ttf] + 80."""
    routes = {"filter_score_kernel": cnt["filter_reads"], "exact_score_kernel": cnt["dense_reads"],
              ("score_positional_warp_kernel" if layout["max_seg_docs"] <= 256 else "score_positional_kernel"): cnt["positional_reads"]}
    kernel = max(routes.items(), key=lambda x: x[1])[0]
    io = int(total_read_bytes) + 80 * cost["reads"]
    if kernel.startswith("score_positional"):
        alg = cost["bytes_csr"]
        parts = {"csr_formula_bytes": cost["bytes_csr"]}
        model = "A_read (SURVEY.md 8d): len + sum_unique_terms[16 + 8 df + 4 ttf] + 80 per read"
    else:
        rows = cost["clauses"] * layout["row_bytes"]
        alg = rows + io
        parts = {"row_bytes": rows, "io_bytes": io, "doc_weight_bytes_l2_resident_not_counted": 4 * n_docs * cost["reads"]}
        model = "clauses x %d-byte rows + read bytes + 80 B results per read" % layout["row_bytes"]
    achieved = alg / kernel_s / 1e9 if kernel_s > 0 else 0.0
    r = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "byte_model": model, "kernel_ms_per_launch": kernel_s * 1e3}
    r.update(parts)
    if traffic:
        r["traffic_frac"] = traffic / kernel_s / 1e9 / peak
    if not kernel.startswith("score_positional"):
        r["csr_formula_bytes_per_launch"] = cost["bytes_csr"]
    if gap_join and gap_join["join_kernel_ms"] / 1e3 > kernel_s:
        # the gap join in front of the scorer is the dominant kernel of this launch (kmer_skips > 0): its roofline leads,
        # the scorer's follows
        g = gap_join_roofline(gap_join, peak, peak_src)
        g["scorer"] = r
        return g
    return r


def secondary_lines(clf, data, o, dev, local, n, peak, peak_src, main_kw, n_docs):
    """The same resident index queried with the reference's other settings (benchmark/scripts/run.py:137-171 sweeps the three
    query algorithms; Configuration.java:37-44 defaults kmer_skips to 5; Configuration.java:169-179 offers bm25):
    device-resident reads, 1 warm-up + 2 timed steps each."""
    import torch
    out = []
    cases = [("CHAIN_PROXIMITY", dict(query_algorithm="CHAIN_PROXIMITY"), 1.0), ("NAIVE_KMER", dict(query_algorithm="NAIVE_KMER"), 1.0),
             ("kmer_skips=5 (Configuration.java default)", dict(kmer_skips=5), 1.0), ("bm25", dict(scoring_algorithm="bm25"), 1.0)]
    sync = torch.cuda.synchronize
    layout = clf.layout()
    for name, kw, frac in cases:
        if all(main_kw.get(k_) == v for k_, v in kw.items()):
            continue
        m = max(1, int(n * frac))
        total = int(data["offsets"][m])
        clf.set_query(shipped_config(device=local, **kw))

        def step():
            clf.classify_device(m, data["d_seq"].data_ptr(), data["d_off"].data_ptr(), total, o["status"].data_ptr(), o["label"].data_ptr(),
                                o["n_hits"].data_ptr(), o["n_top"].data_ptr(), o["top_doc"].data_ptr(), o["top_score"].data_ptr())
        step()
        el, kernel_ms, pipe_ms, _ = timed_steps(step, 2, sync, clf)
        cnt = clf.last_counters()
        cost = clf.read_cost(data["seq_bytes"][:total], data["offsets"][:m + 1])
        roof = route_roofline(cnt, cost, layout, n_docs, total, kernel_ms / 1e3 / 2, peak, peak_src, gap_join=clf.last_gap_join())
        ok = o["status"][:m] == 0
        out.append({"name": name, "reads": m, "value": m * 2 / el, "unit": "reads/s", "ms_per_step": el / 2 * 1e3,
                    "routes": {"filter": cnt["filter_reads"], "exact_dense": cnt["dense_reads"], "positional": cnt["positional_reads"]},
                    "classified_fraction": float(((o["label"][:m] == 2) & ok).sum() / max(int(ok.sum()), 1)),
                    "roofline": {k_: roof[k_] for k_ in ("kernel", "achieved", "frac", "algorithmic_bytes_per_launch", "byte_model",
                                                         "kernel_ms_per_launch", "clauses", "chunks", "matches", "candidates", "verify_kernel_ms_per_launch") if k_ in roof}})
    clf.set_query(shipped_config(device=local, **main_kw))
    return out


def lclassify_line(clf, data, n, labels):
    """File to file on the resident index: LocalClassifier.classify(inputFasta, classifyOutput, summaryOutput)
    (LocalClassifier.java:60-131) through bsp_host_classify_file -- FASTA reader thread, GPU batches, formatter threads, in-order
    writer of the reference's per-read JSON lines.  Files live in /dev/shm (tmpfs) when it exists."""
    import shutil
    import tempfile
    from biospectra_b200 import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="bsp_lc_", dir=base)
    try:
        off = data["offsets"]
        seq = data["seq_bytes"]
        src = data["src"]
        path = os.path.join(d, "reads.fa")
        with open(path, "wb") as f:
            chunk = []
            for i in range(n):
                chunk.append(b">" + synth.header_of(int(src[i])).encode() + b"\n" + seq[off[i]:off[i + 1]].tobytes() + b"\n")
                if len(chunk) == 65536:
                    f.write(b"".join(chunk)); chunk = []
            f.write(b"".join(chunk))
        in_bytes = os.path.getsize(path)
        threads = min(os.cpu_count() or 4, 16)
        res_path, sum_path = os.path.join(d, "reads.fa.result"), os.path.join(d, "reads.fa.result.sum")
        clf.classify_file(path, res_path, sum_path, worker_threads=threads)          # warm-up (page cache, pinned buffers)
        t0 = time.perf_counter()
        clf.classify_file(path, res_path, sum_path, worker_threads=threads)
        el = time.perf_counter() - t0
        with open(sum_path) as f:
            summ = json.load(f)
        try:
            with open(res_path + ".gpu_metrics.json") as f:
                gm = json.load(f)
        except Exception:
            gm = {}
        out_bytes = os.path.getsize(res_path)
        ok = (summ.get("total") == n and summ.get("classified") == labels["classified"] and summ.get("vague") == labels["vague"]
              and summ.get("unknown") == labels["unknown"])
        return {"name": "lclassify file to file (FASTA in, .result / .result.sum out) on the resident index", "reads": n,
                "value": n / el, "unit": "reads/s", "seconds": el, "worker_threads": threads, "input_bytes": in_bytes,
                "output_bytes": out_bytes, "summary": {k_: summ.get(k_) for k_ in ("total", "unknown", "vague", "classified")},
                "summary_matches_device_labels": bool(ok), "files_in": d.split("/")[1],
                "stage_ms": dict(gm.get("stage_ms", {}), gpu_calls_wall=gm.get("classify_call_wall_ms"), gpu_batches=gm.get("gpu_batches"))}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def large_k_line(k, dev, local, peak, peak_src):
    """One k > 12 index (the authors' sweep goes to k = 24, benchmark/scripts/run.py:11): hash dictionary + CSR postings, no dense
    tables.  10 Gbp of 16-mers need 240 GB of postings, so this line uses the 1 Gbp `mid` reference (300 genomes)."""
    import torch
    from biospectra_b200 import Indexer, synth
    wl = dict(WORKLOADS["mid"], reads=100_000)
    lengths = synth.genome_lengths(wl["genomes"], wl["total_bp"], wl["seed"], 1.5e6, 6.5e6)
    ix = Indexer(shipped_config(device=local, kmer_size=k))
    data = synth.build_reference_and_reads_cuda(ix, lengths, wl["reads"], wl["R"], wl["err"], wl["seed"], wl["read_seed"], dev)
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t0 = time.time()
    clf = ix.close_to_classifier(shipped_config(device=local, kmer_size=k))
    t_build = time.time() - t0
    try:
        n = len(data["offsets"]) - 1
        total = int(data["offsets"][-1])
        o = {k_: torch.empty(n, dtype=torch.int32, device=dev) for k_ in ("status", "label", "n_hits", "n_top")}
        o["top_doc"] = torch.empty(n, 10, dtype=torch.int32, device=dev)
        o["top_score"] = torch.empty(n, 10, dtype=torch.float32, device=dev)

        def step():
            clf.classify_device(n, data["d_seq"].data_ptr(), data["d_off"].data_ptr(), total, o["status"].data_ptr(), o["label"].data_ptr(),
                                o["n_hits"].data_ptr(), o["n_top"].data_ptr(), o["top_doc"].data_ptr(), o["top_score"].data_ptr())
        step()
        el, kernel_ms, _, _ = timed_steps(step, 2, torch.cuda.synchronize, clf)
        cnt = clf.last_counters()
        cost = clf.read_cost(data["seq_bytes"], data["offsets"])
        layout = clf.layout()
        st = clf.stats()
        roof = route_roofline(cnt, cost, layout, st["n_docs"], total, kernel_ms / 1e3 / 2, peak, peak_src)
        # k > 12: every (token, segment) costs a probe of the segment's open-addressing dictionary -- at least the 32-byte sector of
        # the key slot -- before any posting is read, and a 16-mer is present in about one segment only: the probes, not the
        # postings of SURVEY 8(d)'s A_read, are what this path fetches
        probes = cost["tokens"] * layout["segments"]
        alg = cost["bytes_csr"] + 32 * probes
        ks = kernel_ms / 1e3 / 2
        roof.update(algorithmic_bytes_per_launch=alg, achieved=alg / ks / 1e9 if ks > 0 else 0.0, frac=(alg / ks / 1e9 / peak) if ks > 0 else 0.0,
                    byte_model="A_read (SURVEY.md 8d) + one 32-byte sector per (token, segment) dictionary probe; latency-bound random access",
                    dictionary_probes_per_s=probes / ks if ks > 0 else 0.0)
        src_doc = 2 * data["src"].astype(np.int64)
        top = o["top_doc"][:, 0].cpu().numpy()
        return {"name": "k=%d (hash dictionary + CSR postings), mid reference: %d genomes / %.1f Gbp" % (k, wl["genomes"], wl["total_bp"] / 1e9),
                "reads": n, "value": n * 2 / el, "unit": "reads/s", "ms_per_step": el / 2 * 1e3, "build_seconds": t_build,
                "index": dict(st, **layout), "source_doc_is_top_hit": float((top == src_doc).mean()),
                "routes": {"filter": cnt["filter_reads"], "exact_dense": cnt["dense_reads"], "positional": cnt["positional_reads"]},
                "roofline": {k_: roof[k_] for k_ in ("kernel", "achieved", "frac", "algorithmic_bytes_per_launch", "byte_model", "kernel_ms_per_launch",
                                                     "dictionary_probes_per_s")}}
    finally:
        clf.close()


def build_record(wl, stats, t_build, k, peak):
    """Index-build byte model (with this build's 8-byte (term, ordinal) tuples): 1 B/base of ASCII in, 0.375 B/base
    of packed reference out and in again; per token (both strand docs) one 8-byte tuple written, ceil(2k/7)-pass LSD radix sort
    (histogram read + scatter read + write = 24 B per pass), CSR: 16 B/token read in two passes, 4 B/position + 8 B/posting
    written; dense tables (k <= 12): 12 B/token read."""
    bases, tokens, postings = wl["total_bp"], stats["n_positions"], stats["n_postings"]
    passes = -(-2 * k // 7) if 2 * k > 16 else -(-2 * k // 8)
    b = bases * (1 + 2 * 0.375) + tokens * (8 + passes * 24 + 16 + 4 + (12 if k <= 12 else 0)) + postings * 8
    return {"seconds": t_build, "gbp_per_s": bases / 1e9 / t_build, "algorithmic_bytes": int(b), "radix_passes": passes,
            "achieved_gbs": b / t_build / 1e9, "frac_of_hbm_peak": b / t_build / 1e9 / peak,
            "note": "wall clock of bsp_index_close_to_classifier (pack -> tuples -> radix sort -> CSR -> dense tables, incl. "
                    "allocations); every replica builds the whole index, so this figure does not grow with --gpus (see docpart.build)"}


def docpart_subrun(args, wl, lengths, data, expected, dev, local, rank, world):
    """N > 1: the same reference split N ways by contiguous genome ranges (doc-partitioned mode, BASELINE configs[3] style), every
    rank scoring the SAME reads (a prefix of rank 0's, broadcast once, resident on the device), hit lists exchanged with ONE NCCL
    all_gather per step and merged.  rank 0 compares the merged results with what its full replica returned for the same reads."""
    import torch
    import torch.distributed as dist
    from biospectra_b200 import Indexer, partition, synth
    m = int(os.environ.get("BSP_DOCPART_READS", 200_000))
    m = min(m, len(data["offsets"]) - 1)
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == 0:
        meta[0], meta[1] = m, int(data["offsets"][m])
    dist.broadcast(meta, 0)
    m, total = int(meta[0]), int(meta[1])
    d_seq = torch.empty(total, dtype=torch.uint8, device=dev)
    d_off = torch.empty(m + 1, dtype=torch.int64, device=dev)
    if rank == 0:
        d_seq.copy_(data["d_seq"][:total])
        d_off.copy_(data["d_off"][:m + 1])
    dist.broadcast(d_seq, 0)
    dist.broadcast(d_off, 0)
    offsets = d_off.cpu().numpy()
    plan = partition.plan_entries(lengths, world)
    e0, e1 = plan[rank]
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.time()
    ix = Indexer(shipped_config(device=local))
    synth.build_reference_and_reads_cuda(ix, lengths, 0, wl["R"], wl["err"], wl["seed"], wl["read_seed"], dev,
                                         rank_slice=lambda g: e0 <= g < e1)
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    clf = ix.close_to_classifier(shipped_config(device=local, query_algorithm=args.algorithm, kmer_skips=args.skips))
    dp = partition.DocPartitionedClassifier(partition.GpuPartitionEngine(clf, dev)).sync_statistics()
    torch.cuda.synchronize()
    tb = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    t_build = float(tb[0])
    dev_reads = (d_seq, d_off, total)

    def step():
        return dp.classify(None, offsets, device_reads=dev_reads)

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    for _ in range(2):
        step()
    c0 = dp.collectives
    steps = 3
    sync()
    t = time.perf_counter()
    for _ in range(steps):
        res = step()
    sync()
    tt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    el = float(tt[0])
    rec = None
    if rank == 0:
        got = {k_: v.cpu().numpy() for k_, v in res.items()}
        same = compare_results(got, expected)
        rec = {"mode": "doc-partitioned x%d: %d genomes / %.2f Gbp split by contiguous genome ranges, %d reads scored by every GPU"
                       % (world, wl["genomes"], wl["total_bp"] / 1e9, m),
               "value": m * steps / el, "unit": "reads/s", "ms_per_step": el / steps * 1e3, "steps": steps, "warmup": 2, "scaling": "strong",
               "collectives_per_step": (dp.collectives - c0) // steps, "collective": "ncclAllGather of %d B per rank (84 B per read)" % (m * 84),
               "plan": plan, "global_docs": dp.global_docs,
               "build": {"seconds_max_over_ranks": t_build, "gbp_per_s": wl["total_bp"] / 1e9 / t_build, "gen_seconds_rank0": t_gen,
                         "note": "every GPU builds only its slice (+ df all_reduce): whole-reference build rate at %d GPUs" % world},
               "verify": {"reads": m, "identical_to_replica": int(same.sum()), "mismatching_reads": np.nonzero(~same)[0][:10].tolist()}}
    clf.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("BSP_WORKLOAD", "C2"))
    ap.add_argument("--algorithm", default="PAIRED_PROXIMITY")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary lines (other query settings, k=16) and the docpart sub-run")
    ap.add_argument("--skips", type=int, default=0, help="kmer_skips of the query side (shipped configuration.json: 0; code default: 5)")
    ap.add_argument("--k", type=int, default=10, help="kmer_size (shipped configuration.json: 10)")
    ap.add_argument("--scoring", default="tfidf", choices=["tfidf", "bm25"])
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads per GPU")
    ap.add_argument("--read-len", type=int, default=0, help="override the mean read length (BASELINE configs[4] sweep: 100-250)")
    ap.add_argument("--err", type=float, default=-1.0, help="override the substitution rate (configs[4] sweep: 0-0.08)")
    ap.add_argument("--mode", default="replica", choices=["replica", "docpart"],
                    help="replica: read-sharded index replicas (default, no collective); docpart: doc-partitioned index")
    ap.add_argument("--verify", action="store_true", help="docpart: compare with a single-index run on rank 0")
    ap.add_argument("--profile", default="classify", choices=["classify", "build"],
                    help="which region cudaProfilerStart/Stop brackets (for ncu --profile-from-start off)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.reads:
        wl["reads"] = args.reads
    if args.read_len:
        wl["R"] = args.read_len
    if args.err >= 0:
        wl["err"] = args.err
    from biospectra_b200 import synth
    lengths = synth.genome_lengths(wl["genomes"], wl["total_bp"], wl["seed"], *((1, 1) if wl.get("equal") else (1.5e6, 6.5e6)))
    if args.impl == "reference":
        return run_reference(args, wl, lengths)
    if args.mode == "docpart":
        return run_docpart(args, wl, lengths)

    import torch
    import torch.distributed as dist
    from biospectra_b200 import Indexer
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- build: synthetic reference generated in HBM, indexed on the GPU (not timed as part of the metric)
    main_kw = dict(query_algorithm=args.algorithm, kmer_skips=args.skips, scoring_algorithm=args.scoring)
    default_cfg = args.k == 10 and args.skips == 0 and args.scoring == "tfidf"
    iconf = shipped_config(device=local, kmer_size=args.k)
    qconf = shipped_config(device=local, kmer_size=args.k, **main_kw)
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline and args.k == 10
    S = sample_plan(wl, lengths)
    t0 = time.time()
    ix = Indexer(iconf)
    # every rank indexes the same reference (replica) but classifies its own reads (read_seed + rank)
    data = synth.build_reference_and_reads_cuda(ix, lengths, wl["reads"], wl["R"], wl["err"], wl["seed"], wl["read_seed"] + rank,
                                                dev, keep_genomes=set(range(S)) if want_cpu else (),
                                                n_frac=float(os.environ.get("BSP_N_READ_FRAC", 0)))
    torch.cuda.empty_cache()        # the generator's scratch goes back to the driver before the index is built
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    if args.profile == "build":
        torch.cuda.cudart().cudaProfilerStart()
    clf = ix.close_to_classifier(qconf)
    if args.profile == "build":
        torch.cuda.cudart().cudaProfilerStop()
    t_build = time.time() - t0
    stats = clf.stats()
    mem = clf.memory_info()
    layout = clf.layout()
    n = len(data["offsets"]) - 1
    total_bytes = int(data["offsets"][-1])
    d_seq, d_off = data["d_seq"], data["d_off"]
    o = {k: torch.empty(n, dtype=torch.int32, device=dev) for k in ("status", "label", "n_hits", "n_top")}
    o["top_doc"] = torch.empty(n, 10, dtype=torch.int32, device=dev)
    o["top_score"] = torch.empty(n, 10, dtype=torch.float32, device=dev)

    def step_device():
        clf.classify_device(n, d_seq.data_ptr(), d_off.data_ptr(), total_bytes, o["status"].data_ptr(), o["label"].data_ptr(),
                            o["n_hits"].data_ptr(), o["n_top"].data_ptr(), o["top_doc"].data_ptr(), o["top_score"].data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        el, kernel_ms, pipe_ms, launches = timed_steps(fn, steps, barrier, clf)
        if world > 1:
            tt = torch.tensor([el, pipe_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el, pipe_ms = float(tt[0]), float(tt[1])
        return el, kernel_ms, pipe_ms, launches

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    # device-resident metric: device time (CUDA events on the library's stream, whole pipeline), max over ranks
    # `ncu --profile-from-start off` captures exactly the timed device-resident steps
    if args.profile == "classify":
        torch.cuda.cudart().cudaProfilerStart()
    el_dev, kernel_ms, pipe_ms, launches = timed(step_device, args.steps)
    if args.profile == "classify":
        torch.cuda.cudart().cudaProfilerStop()
    cnt = clf.last_counters()
    gj = clf.last_gap_join()
    # end to end through the host-buffer call
    host_out = {"r": clf.alloc_outputs(n, pinned=True)}     # caller-provided result arrays, page-locked like the reads

    def step_host():
        clf.classify_packed(data["seq_bytes"], data["offsets"], out=host_out["r"])

    step_host()
    el_e2e, _, _, _ = timed(step_host, args.steps)
    sampler.stop_flag.set()
    sampler.join()
    res = host_out["r"]
    dres = {k: v.cpu().numpy() for k, v in o.items()}
    assert np.array_equal(dres["label"], res["label"]) and np.array_equal(dres["top_doc"], res["top_doc"]), "device/host paths disagree"
    # properties at full size.  tools/verify_simulator_result.py:10-39 expects a classified read's top hit to name its source
    # genome.  With Lucene's tf-idf that only holds while (k+1)-mers are rare.  At C2 they are not: a random 6.5 Mbp genome holds
    # ~79 % of all 11-mers, so it matches ~79 % of a read's pair clauses by chance, while the true source of a read with 4 %
    # substitutions matches the ~64 % of its clauses no substitution touches plus its own chance share of the rest -- a short
    # source genome loses to the long ones.  That is reference semantics: the oracle returns the same hits, bit for bit, on the
    # sample index (cpu_baseline.gpu_parity_on_sample.benchmarked.source_doc_is_top_hit: gpu == oracle).  The decomposition
    # below reports what the argument predicts: when the source is not the top hit, the top hit is usually a longer document.
    ok = res["status"] == 0
    top = res["top_doc"][:, 0]
    src_doc = 2 * data["src"].astype(np.int64)
    hit = (top == src_doc) & ok
    recovered = float(hit.sum() / max(ok.sum(), 1))
    miss = ok & ~hit & (top >= 0)
    doc_len = np.repeat(np.asarray(lengths, np.int64), 2)
    miss_longer = float((doc_len[np.clip(top[miss], 0, len(doc_len) - 1)] > doc_len[src_doc[miss]]).mean()) if miss.any() else None
    labels = {nm: int((res["label"][ok] == i).sum()) for i, nm in enumerate(("unknown", "vague", "classified"))}
    expected = None
    if world > 1 and not args.no_secondary and rank == 0:
        m_dp = min(int(os.environ.get("BSP_DOCPART_READS", 200_000)), n)
        expected = {k: np.array(v[:m_dp]) for k, v in res.items() if not k.startswith("_")}

    total_reads = n * world
    steps = args.steps
    value = total_reads * steps / el_dev
    e2e = total_reads * steps / el_e2e
    line = None
    if rank == 0:
        cost = clf.read_cost(data["seq_bytes"], data["offsets"])
        peak, peak_src = peaks()
        per_launch_s = (kernel_ms / 1e3) / steps
        roof = route_roofline(cnt, cost, layout, stats["n_docs"], total_bytes, per_launch_s, peak, peak_src, gap_join=gj)
        if roof["kernel"] != "gap_join_kernel":          # (the gap join's record carries its own committed capture)
            roof["traffic"] = measured_traffic(args, roof["kernel"]) if default_cfg else None
            if roof["traffic"]:
                roof["traffic_frac"] = roof["traffic"] / per_launch_s / 1e9 / peak
        roof["note"] = ("achieved = algorithmic bytes the launch must fetch from DRAM (byte_model) / the scoring kernel's time from CUDA "
                        "events on the library's stream; traffic = dram bytes of the committed ncu capture of this kernel on this workload")
        line = {"metric": "reads/sec classified", "value": value, "unit": "reads/s", "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": el_dev / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 scores, f64 accumulate", "data": "synthetic", "config": workload_config(args, wl),
                "clocks": sampler.summary(),
                "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(total_bytes + 8 * (n + 1)) * world,
                        "d2h_bytes_per_step": int(n * (4 * 4 + 80) + 4 * n) * world,
                        "note": "bsp_classify_packed with page-locked host buffers: reads + offsets in, status/label/n_hits/n_top/"
                                "top_doc/top_score/routes out, all inside the timed region; bytes are summed over ranks"},
                "gpu_launches": int(launches), "roofline": roof,
                "probes_per_s": cost["unique_terms"] * world * steps / el_dev,
                "device_event_ms_per_step": pipe_ms / steps,
                "timing": "value/e2e: wall clock of exactly K steps between barrier+synchronize, max over ranks; "
                          "device_event_ms_per_step and roofline: CUDA events on the library's stream",
                "index": dict(stats, gen_seconds=t_gen, build_seconds=t_build, build_gbp_per_s=wl["total_bp"] / 1e9 / t_build, **mem, **layout),
                "build": build_record(wl, stats, t_build, args.k, peak),
                "routes": {"filter": cnt["filter_reads"], "exact_dense": cnt["dense_reads"], "positional": cnt["positional_reads"],
                           "reads": cnt["reads"], "filter_rescored_docs_per_read": cnt["filter_candidates"] / max(cnt["filter_reads"], 1),
                           "filter_exception_cells": cnt["filter_exception_cells"], "filter_overflow_reads": cnt["filter_overflow_reads"]},
                "parity_properties": {"source_doc_is_top_hit": recovered, "top_hit_is_longer_doc_when_source_is_not_top": miss_longer,
                                      "labels": labels, "dropped": int((res["status"] == 1).sum()), "errors": int((res["status"] == 2).sum())}}
        if world == 1 and not args.no_secondary and args.k == 10:
            line["secondary"] = secondary_lines(clf, data, o, dev, local, n, peak, peak_src, main_kw, stats["n_docs"])
            if (res["status"] == 0).all():
                line["secondary"].append(lclassify_line(clf, data, n, labels))
    clf.close()
    del o
    torch.cuda.empty_cache()
    if world > 1 and not args.no_secondary:
        rec = docpart_subrun(args, wl, lengths, data, expected, dev, local, rank, world)
        if rank == 0:
            line["docpart"] = rec
    if rank == 0:
        if world == 1 and not args.no_secondary and args.k == 10:
            del data["d_seq"], data["d_off"]
            torch.cuda.empty_cache()
            for k_big in (16, 20):            # benchmark/scripts/run.py:137-171 sweeps k = 5..24
                line["secondary"].append(large_k_line(k_big, dev, local, *peaks()))
                torch.cuda.empty_cache()
        if want_cpu:
            n_cpu = int(os.environ.get("BSP_CPU_SAMPLE_READS", 200000))
            points, arm, (dt, res_o) = cpu_arm_points(wl, lengths, data["kept"], S, n_cpu, ALG_ID[args.algorithm], skips=args.skips,
                                                      scoring=1 if args.scoring == "bm25" else 0)
            rep = cpu_report(points, arm, dt, res_o, stats["n_docs"])
            rep["gpu_parity_on_sample"] = gpu_parity_on_sample(arm, res_o, data["kept"], wl, lengths, local, main_kw)
            line["cpu_baseline"] = rep
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
