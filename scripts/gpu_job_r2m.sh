set -x
mkdir -p gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m/pytest_gpu.log
tail -5 gpurun_out/r2m/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2m/bench_default.json 2> gpurun_out/r2m/bench_default.err; echo "rc=$?" >> gpurun_out/r2m/bench_default.err
tail -3 gpurun_out/r2m/bench_default.err
