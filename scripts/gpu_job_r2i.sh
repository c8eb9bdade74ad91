set -x
mkdir -p gpurun_out/r2i
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section SpeedOfLight --clock-control none --import-source on --profile-from-start off -k regex:gap_join_kernel -c 1 -o gpurun_out/r2i/gap_join_r2i python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2i/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2i/ncu.log
