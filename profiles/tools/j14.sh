#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j14_pytest.log
tail -6 gpurun_out/j14_pytest.log
timeout 300 python profiles/tools/seam_rate.py 20000 > gpurun_out/j14_seam.jsonl 2> gpurun_out/j14_seam.err; cat gpurun_out/j14_seam.jsonl; tail -3 gpurun_out/j14_seam.err
timeout 300 python bench.py --numerics fast --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/j14_bench_fast.json 2> gpurun_out/j14_bench_fast.err; cut -c1-200 gpurun_out/j14_bench_fast.json
