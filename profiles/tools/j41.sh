#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c5.py -m gpu -x -q > gpurun_out/j41_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j41_pytest.log
tail -4 gpurun_out/j41_pytest.log
bash profiles/tools/jv.sh C5_small base old fused ll10 rs12 rs16
cp gpurun_out/jv_variants.txt gpurun_out/j41_variants.txt
timeout 300 python profiles/tools/bvh_bench.py > gpurun_out/j41_bvh_bench.jsonl 2> gpurun_out/j41_bvh_bench.err; tail -3 gpurun_out/j41_bvh_bench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 4 -c 1 -o gpurun_out/j41_bvh -f python bench.py --workload C5_small --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j41_ncu.log 2>&1; ls -la gpurun_out/j41_bvh.ncu-rep
