set -x
mkdir -p gpurun_out/r2c
timeout 600 python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2c/bench_skips5.json 2> gpurun_out/r2c/bench_skips5.err; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gap_join_kernel -c 1 -o gpurun_out/r2c/gap_join_r2c python bench.py --skips 5 --reads 200000 --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2c/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2c/ncu.log
