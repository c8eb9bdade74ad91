set -x
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_gapjoin.py -x -q > gpurun_out/r2b/pytest_gapjoin.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/pytest_gapjoin.log
tail -30 gpurun_out/r2b/pytest_gapjoin.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/pytest_gpu.log
tail -30 gpurun_out/r2b/pytest_gpu.log
timeout 600 python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2b/bench_skips5.json 2> gpurun_out/r2b/bench_skips5.err; echo "rc=$?"
tail -5 gpurun_out/r2b/bench_skips5.err
