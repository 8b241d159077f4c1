#!/bin/bash
# round-2 job 1: parity holes + baselines + fast-numerics variant + seam rate
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/j1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j1_bench_exact.json 2> gpurun_out/j1_bench_exact.err
timeout 300 python profiles/tools/bench_variant.py ada-ray-tracer_b200/csrc/_variants/libgcore_fast.so --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j1_bench_fast.json 2> gpurun_out/j1_bench_fast.err
timeout 300 python profiles/tools/seam_rate.py 20000 > gpurun_out/j1_seam.jsonl 2> gpurun_out/j1_seam.err
timeout 600 python bench.py --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j1_bench_c5small.json 2> gpurun_out/j1_bench_c5small.err
tail -3 gpurun_out/j1_pytest.log; cat gpurun_out/j1_bench_exact.json gpurun_out/j1_bench_fast.json gpurun_out/j1_seam.jsonl gpurun_out/j1_bench_c5small.json | cut -c1-600
