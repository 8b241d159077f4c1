#!/bin/bash
# round-2 job 2: refactored library (per-device state, committed-scene render, queue-based LBVH traversal)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j2_bench_exact.json 2> gpurun_out/j2_bench_exact.err
timeout 600 python bench.py --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j2_bench_c5small.json 2> gpurun_out/j2_bench_c5small.err
tail -5 gpurun_out/j2_pytest.log; cat gpurun_out/j2_bench_exact.json gpurun_out/j2_bench_c5small.json | cut -c1-300; tail -3 gpurun_out/j2_bench_c5small.err
