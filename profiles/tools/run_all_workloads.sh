#!/bin/bash
# every BASELINE config on 1 GPU (bench.py --workload ...), one JSON line each -> gpurun_out/r02_workloads_1gpu.jsonl
out=gpurun_out/r02_workloads_1gpu.jsonl; : > $out
for w in C1 C2 C3 C4_1024 C4 C5; do
  extra=""; [ $w = C4_1024 ] || extra="--no-secondary"
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 $extra 2>gpurun_out/wl_$w.err | tail -n 1 >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/r02_workloads_1gpu.jsonl"):
    d = json.loads(l)
    cb = d.get("cpu_baseline") or {}
    print("%-8s %9.2f ms/step  %10.1f Mrays/s  %9.1f Msamples/s  e2e %10.1f  cpu %8.1f (%s cores)  roof %.4f" % (
        d["config"]["workload"][:8], d["ms_per_step"], d["value"], d["msamples_per_s"], d["e2e"]["value"], cb.get("value", 0), cb.get("cores"), d["roofline"]["frac"]))
PY
