#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/j6_variants.txt
for n in base i j k l m n o; do
  if [ $n = base ]; then lib=ada-ray-tracer_b200/libgcore_b200.so; else lib=ada-ray-tracer_b200/csrc/_variants/libgcore_$n.so; fi
  timeout 300 python profiles/tools/bench_variant.py $lib --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j6_$n.json 2> gpurun_out/j6_$n.err
  python - $n <<'PY' >> gpurun_out/j6_variants.txt
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/j6_%s.json" % n))
    print(n, "%.2f ms/step  %.0f Mrays/s  nodes/trav %.1f" % (d["ms_per_step"], d["value"], d["roofline"]["nodes_per_traversal"]))
except Exception as e:
    print(n, "failed", e)
PY
done
cat gpurun_out/j6_variants.txt
