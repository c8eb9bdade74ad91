set -x
mkdir -p gpurun_out/r2g
BSP_LIB=$PWD/_variants/lib_bm2048.so timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section SchedulerStats --section InstructionStats --clock-control none --import-source on --profile-from-start off -k regex:gap_join_kernel -c 1 -o gpurun_out/r2g/gap_join_r2g python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2g/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2g/ncu.log
