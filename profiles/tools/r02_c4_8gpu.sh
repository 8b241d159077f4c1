#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 5 --warmup 3 --workload C4 > gpurun_out/r02f_bench_c4_8gpu.json 2> gpurun_out/r02f_bench_c4_8gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_bench_c4_8gpu.json").read().strip().splitlines()[-1])
print("C4 x8: %.3f ms/step  %.1f Mrays/s  e2e %.1f  per-rank kernel ms %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], ["%.2f" % x for x in d.get("kernel_ms_per_rank", [])]))
PY
