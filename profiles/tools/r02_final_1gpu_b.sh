#!/bin/bash
# after the last kernel change (no spill in the traversal warps): tests again, every workload again, launch list and both full captures again
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_1gpu.log
tail -3 gpurun_out/r02f_pytest_1gpu.log
bash profiles/tools/run_all_workloads.sh > gpurun_out/r02f_workloads.txt 2>&1; cat gpurun_out/r02f_workloads.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err
timeout 300 python profiles/tools/bvh_bench.py > gpurun_out/r02f_bvh_bench.jsonl 2> gpurun_out/r02f_bvh_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 4 -c 1 -o gpurun_out/r02f_wave -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_ncu_wave.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bvh_kernel -s 4 -c 1 -o gpurun_out/r02f_bvh -f python bench.py --workload C5_small --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bvh.log 2>&1
ls -la gpurun_out/r02f_*
