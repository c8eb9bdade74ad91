set -x
mkdir -p gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_gapjoin.py tests/test_reference_reads.py tests/test_gpu_lucene_golden.py -x -q > gpurun_out/r2s/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s/pytest.log
tail -4 gpurun_out/r2s/pytest.log
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on --profile-from-start off -k regex:gap_join_kernel -c 1 -o gpurun_out/r2s/gap_join_r2s python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2s/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2s/ncu.log
