#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_holes.py tests/test_gpu_bvh.py tests/test_gpu_committed.py -m gpu -x -q > gpurun_out/j56_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j56_pytest.log
tail -5 gpurun_out/j56_pytest.log
timeout 300 python profiles/tools/seam_rate.py > gpurun_out/j56_seam.jsonl 2> gpurun_out/j56_seam.err; cut -c1-330 gpurun_out/j56_seam.jsonl; tail -3 gpurun_out/j56_seam.err
