#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_c5.py -m gpu -x -q > gpurun_out/j44_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j44_pytest.log
tail -15 gpurun_out/j44_pytest.log
bash profiles/tools/jv.sh C5_small base nogrid base nogrid
cp gpurun_out/jv_variants.txt gpurun_out/j44_variants.txt
