set -x
mkdir -p gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t/pytest_gpu.log
tail -5 gpurun_out/r2t/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2t/smoke.log; tail -2 gpurun_out/r2t/smoke.log
timeout 1200 python bench.py > gpurun_out/r2t/bench_default.json 2> gpurun_out/r2t/bench_default.err; echo "rc=$?" >> gpurun_out/r2t/bench_default.err
tail -3 gpurun_out/r2t/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2t/bench_ref.json 2> gpurun_out/r2t/bench_ref.err
