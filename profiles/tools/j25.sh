#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j25_pytest.log
tail -5 gpurun_out/j25_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/j25_bench.json 2> gpurun_out/j25_bench.err; cut -c1-230 gpurun_out/j25_bench.json
GCORE_TAIL_ITEMS=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/j25_bench_notail.json 2> gpurun_out/j25_bench_notail.err; cut -c1-230 gpurun_out/j25_bench_notail.json
timeout 300 python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/j25_bench_c3.json 2> gpurun_out/j25_bench_c3.err; cut -c1-230 gpurun_out/j25_bench_c3.json
GCORE_TAIL_ITEMS=0 timeout 300 python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/j25_bench_c3_notail.json 2> gpurun_out/j25_bench_c3_notail.err; cut -c1-230 gpurun_out/j25_bench_c3_notail.json
