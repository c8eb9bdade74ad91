set -x
mkdir -p gpurun_out/r2e
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gap_join_kernel -c 1 -o gpurun_out/r2e/gap_join_r2e python bench.py --skips 5 --reads 200000 --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2e/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2e/ncu.log
