#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_c5.py -m gpu -x -q > gpurun_out/j42_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j42_pytest.log
tail -4 gpurun_out/j42_pytest.log
bash profiles/tools/jv.sh C5_small base spec spec10 t256s320 t224s352 t320s256 t192s384 ll8 ll12rs12
cp gpurun_out/jv_variants.txt gpurun_out/j42_variants.txt
