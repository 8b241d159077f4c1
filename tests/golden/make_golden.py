#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/art_oracle.cpp).

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own and cannot be run here (Ada, no GNAT), so these
are outputs of the ORACLE, which tests/test_oracle_kat.py pins independently (numpy restatement of the Ada sources,
Random123 KATs, closed-form radiometry).  They serve two purposes: the -m gpu tests can check the CUDA path against
committed numbers without trusting the oracle build of the GPU box, and the CPU suite detects any drift of the oracle
(compiler, flags, host CPU).  Inputs are seeded; every fixture stores its inputs next to its outputs.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402

RENDER_CASES = [
    # name, render_type, aa, depth, units, w, h, seed
    ("mis_d8", "PT_MIS", 1, 8, 2, 48, 32, 0x5EED0004),
    ("mis_noaa", "PT_MIS", 0, 8, 3, 40, 24, 7),
    ("shadow_d1", "PT_SHADOW", 1, 1, 4, 48, 32, 0x5EED0003),      # BASELINE config C3 shape
    ("shadow_d8", "PT_SHADOW", 1, 8, 1, 40, 24, 11),
    ("stupid_d8", "PT_STUPID", 1, 8, 2, 40, 24, 13),
]


def random_rays(rng, n):
    rays = np.zeros(n, ob.ray_dt)
    rays["origin"] = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 12.0], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 8
    d[:k] = 0; d[np.arange(k), rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)
    rays["direction"] = d.astype(np.float32)
    return rays


def soup(rng, ntris, extent=2.0, size=0.3):
    c = rng.uniform(-extent, extent, (ntris, 1, 3))
    v = (c + rng.uniform(-size, size, (ntris, 3, 3))).astype(np.float32).reshape(-1, 3)
    return v, np.arange(3 * ntris, dtype=np.int32).reshape(-1, 3)


def main():
    osc = ob.OracleScene()
    meta = {}
    # ---- C2: primary hit table at 1024x1024 (44 B x 1 Mi records): digest + tallies + a 1/256 sample
    hits = osc.primary_hits(1024, 1024)
    meta["c2_sha256"] = hashlib.sha256(hits.tobytes()).hexdigest()
    meta["c2_hit_pixels"] = int(hits["is_hit"].sum())
    meta["c2_by_prim_type"] = {str(k): int(((hits["prim_type"] == k) & (hits["is_hit"] == 1)).sum()) for k in range(4)}
    meta["c2_triangle_ids"] = sorted(int(x) for x in np.unique(hits["prim_index"][(hits["prim_type"] == 2) & (hits["is_hit"] == 1)]))
    hits768 = osc.primary_hits(1024, 768)
    meta["c1_primary_sha256"] = hashlib.sha256(hits768.tobytes()).hexdigest()
    np.save(os.path.join(HERE, "c2_primary_1024_every16.npy"), hits[::16, ::16].copy())
    # ---- scene queries on seeded random rays
    rng = np.random.default_rng(20261018)
    rays = random_rays(rng, 4096)
    np.savez_compressed(os.path.join(HERE, "closest_random.npz"), rays=rays, hits=osc.closest_hit_batch(rays))
    # ---- small renders: HDR accumulators, ray counts, LDR frames
    out = {}
    for name, rtype, aa, depth, units, w, h, seed in RENDER_CASES:
        acc, nrays = osc.render(ob.make_params(w, h, getattr(ob, rtype), aa=aa, max_depth=depth, passes=units, seed=seed, threads=1))
        spp = units * (4 if aa else 1)
        out[name + "_acc"] = acc
        out[name + "_ldr"] = ob.resolve(acc, spp)
        meta["render_" + name] = {"render_type": rtype, "aa": aa, "depth": depth, "units": units, "w": w, "h": h, "seed": seed,
                                  "rays": nrays, "spp": spp}
    np.savez_compressed(os.path.join(HERE, "renders_small.npz"), **out)
    # ---- triangle-only nearest-hit queries (legacy gcore seam): soup of 500 triangles, 3000 rays
    verts, tris = soup(rng, 500)
    r2 = np.zeros(3000, ob.ray_dt)
    r2["origin"] = rng.uniform(-3, 3, (3000, 3)).astype(np.float32)
    d = rng.normal(size=(3000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r2["direction"] = d.astype(np.float32)
    prim = np.zeros(3000, np.int32); t = np.zeros(3000, np.float32); u = np.zeros(3000, np.float32); v = np.zeros(3000, np.float32)
    ob.lib().ao_bruteforce_closest(verts.ctypes.data, tris.ctypes.data, len(tris), 3000, r2.ctypes.data, 0.0, 100000.0,
                                   prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
    np.savez_compressed(os.path.join(HERE, "soup500.npz"), verts=verts, tris=tris, rays=r2, prim=prim, t=t, u=u, v=v)
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
