"""bench.py's reference arm runs on the CPU (the oracle on all host cores) and must print ONE JSON line with the keys the
driver reads; the b200 arm needs a GPU and is exercised by the driver itself."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C3", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1024x1024" in cb["sample"]
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_recorded_ncu_traffic_belongs_to_this_source():
    """roofline.traffic in the bench line comes from profiles/r02_traffic.json (one `ncu --set full` capture per render kernel).  The
    file names the hash of the kernel sources it was taken on; bench.py reports null when that is not the source it runs, and this test
    fails so that a kernel change cannot ship with a stale capture (re-run profiles/tools/r02_final_1gpu_b.sh, or delete the file)."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        pytest.skip("no capture recorded: roofline.traffic is null")
    tj = json.load(open(path))
    assert tj["kernel_source_sha16"] == bench.kernel_source_hash()
    for wl in ("C4_1024", "C5_small"):
        got = bench.recorded_traffic(wl)
        assert got[0] == tj["workloads"][wl]["dram_bytes_per_launch"] > 0 and got[1] is None
