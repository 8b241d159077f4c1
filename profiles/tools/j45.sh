#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_holes.py tests/test_gpu_bvh.py tests/test_gpu_committed.py -m gpu -x -q > gpurun_out/j45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j45_pytest.log
tail -15 gpurun_out/j45_pytest.log
timeout 300 python profiles/tools/seam_rate.py > gpurun_out/j45_seam.jsonl 2> gpurun_out/j45_seam.err; cat gpurun_out/j45_seam.jsonl; tail -3 gpurun_out/j45_seam.err
GCORE_SEAM_SERVER=0 timeout 300 python profiles/tools/seam_rate.py > gpurun_out/j45_seam_launch.jsonl 2> gpurun_out/j45_seam_launch.err; cat gpurun_out/j45_seam_launch.jsonl
