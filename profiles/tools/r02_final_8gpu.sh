#!/bin/bash
# round 2, final source, one 8-GPU box: the default bench at N = 8 and 4 (and 2), C5 at N = 8, the C harness replaying test.adb on 8 GPUs
mkdir -p gpurun_out
run() { # N workload tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 5 --warmup 3 --workload $2 > gpurun_out/r02f_bench_$3.json 2> gpurun_out/r02f_bench_$3.err
  python - $3 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02f_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.3f ms/step  %.1f Mrays/s  e2e %.1f  per-rank kernel ms %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], ["%.2f" % x for x in d.get("kernel_ms_per_rank", [])]))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02f_bench_%s.err" % sys.argv[1]).read()[-400:])
PY
}
run 8 C4_1024 8gpu
run 8 C5 c5_8gpu
run 4 C4_1024 4gpu
run 2 C4_1024 2gpu
timeout 300 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/r02f_pytest_8gpu.log 2>&1; tail -2 gpurun_out/r02f_pytest_8gpu.log
