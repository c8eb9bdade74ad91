#!/bin/bash
# The GPU-side evidence of a round in one gpurun call (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash scripts/gpu_evidence.sh r2x'
# tests, smoke, the default bench (+ reference arm), launch lists of the shipped and the kmer_skips = 5 step, and one
# `--set full` capture of the dominant kernel.  Everything lands in gpurun_out/<tag>/; copy what is to be judged to profiles/.
set -x
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?" >> $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_C2_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_launch.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_C2_skips5_step.csv python bench.py --skips 5 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/ncu_launch5.log 2>&1
timeout 1200 $NCU --set full --import-source on -k regex:filter_score_kernel -c 1 -o $OUT/filter_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_full.log 2>&1
tail -n 3 $OUT/pytest_gpu.log; tail -n 3 $OUT/smoke.log
