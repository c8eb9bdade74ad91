set -x
mkdir -p gpurun_out/r2l
B="python bench.py --skips 5 --no-secondary --no-cpu-baseline --steps 2 --warmup 1"
for v in w4m5bm1024 w4m4 w2m8; do
  BSP_LIB=$PWD/_variants/lib_$v.so timeout 600 $B > gpurun_out/r2l/bench_$v.json 2> gpurun_out/r2l/bench_$v.err
done
for f in gpurun_out/r2l/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "reads/s %.0f step %.1f ms join %.1f ms frac %.3f" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_launch",0), r["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
