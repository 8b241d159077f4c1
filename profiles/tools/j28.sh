#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j28_pytest.log
tail -4 gpurun_out/j28_pytest.log
bash profiles/tools/jv.sh C4_1024 base a
bash profiles/tools/jv.sh C4_1024_eighth base a
bash profiles/tools/jv.sh C5_small base a
