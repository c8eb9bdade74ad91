set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a/gpu.txt
nproc >> gpurun_out/r2a/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2a/bench_default.json 2> gpurun_out/r2a/bench_default.err; echo "rc=$?" >> gpurun_out/r2a/bench_default.err
BSP_FILTER_V3=1 timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2a/bench_v3.json 2> gpurun_out/r2a/bench_v3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2a/launches_C2_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2a/ncu_launch.log 2>&1
tail -3 gpurun_out/r2a/pytest_gpu.log
