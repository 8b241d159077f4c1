#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/tools/experiments/exp_c5_spheres.py > gpurun_out/j4_spheres.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bvh_batch -s 2 -c 1 -o gpurun_out/j4_bvh_batch -f python profiles/tools/bvh_bench.py 913 4194304 > gpurun_out/j4_ncu_bvh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 4 -c 1 -o gpurun_out/j4_c5small -f python bench.py --workload C5_small --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j4_ncu_c5.log 2>&1
cat gpurun_out/j4_spheres.txt; ls -la gpurun_out/*.ncu-rep
