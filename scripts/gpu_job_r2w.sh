set -x
mkdir -p gpurun_out/r2w
timeout 600 python -m pytest tests/test_gpu_gapjoin.py -x -q > gpurun_out/r2w/pytest_gapjoin.log 2>&1; echo "rc=$?" >> gpurun_out/r2w/pytest_gapjoin.log
tail -3 gpurun_out/r2w/pytest_gapjoin.log
B="python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 2 --warmup 1"
timeout 600 $B > gpurun_out/r2w/bench_default.json 2> gpurun_out/r2w/bench_default.err
BSP_LIB=$PWD/_variants/lib_sel.so timeout 600 $B > gpurun_out/r2w/bench_sel.json 2> gpurun_out/r2w/bench_sel.err
for f in gpurun_out/r2w/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "reads/s %.0f step %.1f ms join %.1f ms verify %.1f ms frac %.3f" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_launch",0), r.get("verify_kernel_ms_per_launch") or 0, r["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
