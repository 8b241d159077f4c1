#!/bin/bash
mkdir -p gpurun_out
# the decoupled LBVH kernel first, on its own and under a short timeout (spin loops inside)
timeout 300 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_holes.py -x -q -m gpu > gpurun_out/j8_pytest_bvh.log 2>&1; echo "rc=$?" >> gpurun_out/j8_pytest_bvh.log
tail -4 gpurun_out/j8_pytest_bvh.log
if grep -q "rc=0" gpurun_out/j8_pytest_bvh.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j8_pytest.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j8_bench_exact.json 2> gpurun_out/j8_bench_exact.err
  timeout 600 python bench.py --workload C5_small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j8_bench_c5small.json 2> gpurun_out/j8_bench_c5small.err
  tail -4 gpurun_out/j8_pytest.log; cat gpurun_out/j8_bench_exact.json gpurun_out/j8_bench_c5small.json | cut -c1-260
fi
