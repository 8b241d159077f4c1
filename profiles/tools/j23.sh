#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_holes.py tests/test_gpu_committed.py -x -q -m gpu > gpurun_out/j23_pytest_bvh.log 2>&1; echo "rc=$?" >> gpurun_out/j23_pytest_bvh.log
tail -4 gpurun_out/j23_pytest_bvh.log
if grep -q "rc=0" gpurun_out/j23_pytest_bvh.log; then
  timeout 600 python -m pytest tests/test_gpu_c5.py -x -q -m gpu 2>&1 | tail -3
  timeout 600 python bench.py --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j23_bench_c5small.json 2> gpurun_out/j23_bench_c5small.err
  cut -c1-260 gpurun_out/j23_bench_c5small.json
fi
