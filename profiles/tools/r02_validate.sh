#!/bin/bash
# last look at the committed source: smoke(), the whole GPU suite, the race soak, one default bench line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02v_smoke.log 2>&1; tail -1 gpurun_out/r02v_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log; tail -3 gpurun_out/r02v_pytest.log
timeout 600 python profiles/tools/experiments/soak_determinism.py > gpurun_out/r02v_soak.txt 2>&1; tail -2 gpurun_out/r02v_soak.txt
timeout 600 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['secondary']['ms_per_step'], d['secondary']['roofline']['frac'], d['secondary']['roofline']['traffic'], d['gpu_launches'])"
