#!/usr/bin/env python
"""Rays per second through the legacy one-ray seam gcore_closest_hit (scene_hydra_embree.adb:426-446 calls it once per
ray from every Ada task): 1 and 16 host threads, C callers (tests/c_harness/seam_hammer.c) so that no Python sits in the loop."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exe = os.path.join(ROOT, "gpurun_out", "seam_hammer")
os.makedirs(os.path.dirname(exe), exist_ok=True)
subprocess.run(["gcc", "-O2", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "c_harness", "seam_hammer.c"), "-ldl", "-lm"], check=True)
lib = os.path.join(ROOT, "ada-ray-tracer_b200", "libgcore_b200.so")
out = []
for threads in (1, 4, 16, 28):
    r = subprocess.run([exe, lib, str(threads), sys.argv[1] if len(sys.argv) > 1 else "20000"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print(r.stdout, r.stderr); raise SystemExit(1)
    out.append(json.loads(r.stdout.strip().splitlines()[-1]))
    print(json.dumps(out[-1]), flush=True)
