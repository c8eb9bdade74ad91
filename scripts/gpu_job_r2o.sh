set -x
mkdir -p gpurun_out/r2o
B="python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3"
BSP_FILTER_V3=1 timeout 600 $B > gpurun_out/r2o/bench_v3_ldgsts.json 2> gpurun_out/r2o/bench_v3_ldgsts.err
BSP_FILTER_V3=1 BSP_LIB=$PWD/_variants/lib_f3bulk.so timeout 600 $B > gpurun_out/r2o/bench_v3_bulk.json 2> gpurun_out/r2o/bench_v3_bulk.err
timeout 600 $B > gpurun_out/r2o/bench_v1.json 2> gpurun_out/r2o/bench_v1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2o/launches_C2_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2o/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2o/launches_C2_skips5_step.csv python bench.py --skips 5 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2o/ncu_launch5.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:filter_score_kernel -c 1 -o gpurun_out/r2o/filter_r2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2o/ncu_full.log 2>&1
for f in gpurun_out/r2o/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "reads/s %.0f step %.2f ms kernel %.2f ms frac %.3f" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_launch",0), r["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
