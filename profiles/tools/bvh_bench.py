#!/usr/bin/env python
"""K4 micro-benchmark: LBVH closest-hit traversal of the C5 mesh (10 002 828 triangles) for coherent (primary) and
incoherent (random, bounce-like) rays: Mrays/s, nodes / triangles per ray, algorithmic bytes per ray (64 B per node
visited + 48 B per triangle slot tested + 24 B ray in + 40 B hit out) and the fraction of the measured HBM copy
bandwidth that traffic would need.  Prints one JSON line per ray set.   usage: bvh_bench.py [nseg=913] [nrays=4194304]"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 913
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
pkg = ge.load_package()
syn = importlib.import_module(pkg.__name__ + ".synthetic")
pkg.set_device(0)
sc = pkg.Scene(); sc.Init()
pos, nrm, tris = syn.cube_sphere(nseg)
sc.set_mesh(pos, nrm, tris, mesh_mode=pkg.MESH_NEAREST)
info = pkg.GeometryCore().bvh_info()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)

rng = np.random.default_rng(1)
sets = {}
# coherent: camera rays of a sqrt(n) x sqrt(n) frame (camera of scene.adb:212-215, pixel centres), row-major
w = int(np.sqrt(n)); h = n // w
y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
d = np.stack([x + 0.5 - w / 2.0, y + 0.5 - h / 2.0, np.full(x.shape, -float(w))], -1).reshape(-1, 3)
d /= np.linalg.norm(d, axis=1, keepdims=True)
r = np.zeros(len(d), pkg.ray_dt); r["origin"] = (0.0, 2.55, 12.5); r["direction"] = d.astype(np.float32)
sets["primary_%dx%d" % (w, h)] = r
# incoherent: origins uniform in the box, aimed at a Gaussian cloud around the mesh
o = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 4.9], (n, 3))
t = np.array([0.0, 1.6, 2.5]) + rng.normal(size=(n, 3)) * 1.0
d = t - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
r = np.zeros(n, pkg.ray_dt); r["origin"] = o.astype(np.float32); r["direction"] = d.astype(np.float32)
sets["incoherent_box"] = r
for name, rays in sets.items():
    res = pkg.bvh_bench(rays, which=1, repeat=5)
    res.update({"rays": name, "n": len(rays), "ntris": info["ntris"], "bvh_bytes": info["bytes"], "build_ms": info["build_ms"],
                "algorithmic_GBs": res["mrays_s"] * 1e6 * res["bytes_per_ray"] / 1e9, "hbm_peak_GBs": hbm,
                "frac_of_hbm": res["mrays_s"] * 1e6 * res["bytes_per_ray"] / 1e9 / hbm})
    print(json.dumps(res), flush=True)
