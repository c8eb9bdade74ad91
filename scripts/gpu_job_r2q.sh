set -x
mkdir -p gpurun_out/r2r
B="python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 2 --warmup 1"
for v in w9m3 w10m3 w12m2 w16m2; do
  BSP_LIB=$PWD/_variants/lib_$v.so timeout 600 $B > gpurun_out/r2r/bench_$v.json 2> gpurun_out/r2r/bench_$v.err
done
for f in gpurun_out/r2r/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "reads/s %.0f step %.1f ms join %.1f ms verify %.1f ms frac %.3f cands %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_launch",0), r.get("verify_kernel_ms_per_launch") or 0, r["frac"], r.get("candidates")))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
