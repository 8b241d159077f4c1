#!/bin/bash
# usage: jv.sh <workload> <variant names...>   (variant "base" = the product library)
mkdir -p gpurun_out
wl=$1; shift
: > gpurun_out/jv_variants.txt
for n in "$@"; do
  if [ $n = base ]; then lib=ada-ray-tracer_b200/libgcore_b200.so; else lib=ada-ray-tracer_b200/csrc/_variants/libgcore_$n.so; fi
  timeout 300 python profiles/tools/bench_variant.py $lib --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/jv_$n.json 2> gpurun_out/jv_$n.err
  python - $n <<'PY' >> gpurun_out/jv_variants.txt
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/jv_%s.json" % n))
    print(n, "%.2f ms/step  %.0f Mrays/s  e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print(n, "failed", e, open("gpurun_out/jv_%s.err" % n).read()[-300:])
PY
done
cat gpurun_out/jv_variants.txt
