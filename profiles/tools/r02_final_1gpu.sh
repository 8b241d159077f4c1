#!/bin/bash
# round 2, final source: the whole GPU suite, every BASELINE workload on one GPU, the stand-alone traversal rate, the race soak, the ncu
# launch list of the default bench and one full capture of each render kernel.  Outputs under gpurun_out/r02f_*.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
bash profiles/tools/run_all_workloads.sh > gpurun_out/r02f_workloads.txt 2>&1; cat gpurun_out/r02f_workloads.txt
timeout 300 python profiles/tools/bvh_bench.py > gpurun_out/r02f_bvh_bench.jsonl 2> gpurun_out/r02f_bvh_bench.err
timeout 600 python profiles/tools/experiments/soak_determinism.py > gpurun_out/r02f_soak.txt 2>&1; tail -2 gpurun_out/r02f_soak.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; tail -c 600 gpurun_out/r02f_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 4 -c 1 -o gpurun_out/r02f_wave -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_ncu_wave.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bvh_kernel -s 4 -c 1 -o gpurun_out/r02f_bvh -f python bench.py --workload C5_small --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bvh.log 2>&1
ls -la gpurun_out/r02f_*
