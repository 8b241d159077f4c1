#!/bin/bash
# round 2: postponed leaves + fused slab test in the traversal warps -- parity first, then A/B on C5_small
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c5.py tests/test_gpu_holes.py tests/test_gpu_committed.py -m gpu -x -q > gpurun_out/j40_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j40_pytest.log
tail -4 gpurun_out/j40_pytest.log
bash profiles/tools/jv.sh C5_small base old fused post ll3 ll10 ll16 rs4 rs16
cp gpurun_out/jv_variants.txt gpurun_out/j40_variants.txt
