#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/j3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j3_pytest.log
timeout 600 python bench.py --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j3_bench_c5small.json 2> gpurun_out/j3_bench_c5small.err
timeout 600 python profiles/tools/bvh_bench.py > gpurun_out/j3_bvh_bench.jsonl 2> gpurun_out/j3_bvh_bench.err
tail -5 gpurun_out/j3_pytest.log; cat gpurun_out/j3_bench_c5small.json | cut -c1-300; cat gpurun_out/j3_bvh_bench.jsonl
