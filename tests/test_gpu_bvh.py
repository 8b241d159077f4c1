"""GPU LBVH build + traversal (K3/K4) against the oracle: legacy gcore_* seam (the Embree replacement) and the
nearest-hit mesh mode of the renderer.  Primitive ids AND t/u/v must be bit-exact: the boxes only cull."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _syn(pkg):
    spec = importlib.util.spec_from_file_location("ada_ray_tracer_b200_synthetic", os.path.join(os.path.dirname(pkg.__file__), "synthetic.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def _bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy(); b[np.isnan(a)] = 0x7FC00000
    return b


def _soup(rng, ntris, extent=2.0, size=0.3):
    c = rng.uniform(-extent, extent, (ntris, 1, 3))
    v = (c + rng.uniform(-size, size, (ntris, 3, 3))).astype(np.float32).reshape(-1, 3)
    return v, np.arange(3 * ntris, dtype=np.int32).reshape(-1, 3)


def _rays(rng, n, pkg, extent=3.0):
    r = np.zeros(n, pkg.ray_dt)
    r["origin"] = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 16] = 0; d[np.arange(n // 16), rng.integers(0, 3, n // 16)] = 1.0       # axis-parallel: 1/0 = inf in the slab tests
    r["direction"] = d.astype(np.float32)
    return r


def _brute(ob, verts, tris, rays, t0, t1):
    n = len(rays)
    prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    rr = np.ascontiguousarray(rays.view(ob.ray_dt))
    ob.lib().ao_bruteforce_closest(verts.ctypes.data, tris.ctypes.data, len(tris), n, rr.ctypes.data, t0, t1,
                                   prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
    return prim, t, u, v


def _check(found, hits, prim, t, u, v):
    assert np.array_equal(found != 0, prim >= 0)
    m = prim >= 0
    assert np.array_equal(hits["primIndex"][m], prim[m])
    assert np.array_equal(_bits(hits["t"][m]), _bits(t[m]))
    # texCoord = (weight of B, weight of C) = (reference v, reference u)
    assert np.array_equal(_bits(hits["texCoord"][m, 0]), _bits(v[m]))
    assert np.array_equal(_bits(hits["texCoord"][m, 1]), _bits(u[m]))
    assert np.all(hits["primIndex"][~m] == -1)                         # records are only written on a hit


def _check_ng(hits, prim, verts, tris):
    """HitCpp.normal = unnormalised geometric normal (B - A) x (C - A) of the hit triangle, like Embree's Ng"""
    m = prim >= 0
    A, B, C = verts[tris[prim[m], 0]], verts[tris[prim[m], 1]], verts[tris[prim[m], 2]]
    e1, e2 = (B - A).astype(np.float32), (C - A).astype(np.float32)
    ng = np.stack([e1[:, 1] * e2[:, 2] - e2[:, 1] * e1[:, 2], e1[:, 2] * e2[:, 0] - e2[:, 2] * e1[:, 0], e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]], 1)
    assert np.array_equal(_bits(hits["normal"][m]), _bits(ng))


@pytest.mark.parametrize("ntris", [1, 2, 3, 17, 2000])
def test_legacy_seam_random_soup(pkg, ob, ntris):
    rng = np.random.default_rng(100 + ntris)
    verts, tris = _soup(rng, ntris)
    gc = pkg.GeometryCore()
    gc.init_and_clear()
    assert gc.add_mesh_3f(verts, tris) == 0
    gc.commit_scene()
    rays = _rays(rng, 20000, pkg)
    found, hits = gc.closest_hit_batch(rays, 0.0, 100000.0)
    prim, t, u, v = _brute(ob, verts, tris, rays, 0.0, 100000.0)
    _check(found, hits, prim, t, u, v)
    _check_ng(hits, prim, verts, tris)
    if ntris >= 17:
        assert (prim >= 0).sum() > 50
    info = gc.bvh_info()
    assert info["ntris"] == ntris and info["nnodes"] == max(ntris - 1, 0)
    gc.destroy()


def test_legacy_seam_t_window_and_single_ray(pkg, ob):
    rng = np.random.default_rng(5)
    verts, tris = _soup(rng, 500)
    gc = pkg.GeometryCore(); gc.init_and_clear(); gc.add_mesh_3f(verts, tris); gc.commit_scene()
    rays = _rays(rng, 5000, pkg)
    found, hits = gc.closest_hit_batch(rays, 0.5, 2.5)
    prim, t, u, v = _brute(ob, verts, tris, rays, 0.5, 2.5)
    _check(found, hits, prim, t, u, v)
    # one ray per call (the signature scene_hydra_embree.adb:426-446 uses): same answers, record untouched on a miss
    idx = np.concatenate([np.flatnonzero(prim >= 0)[:20], np.flatnonzero(prim < 0)[:20]])
    for i in idx:
        ok, h = gc.closest_hit(rays["origin"][i], rays["direction"][i], 0.5, 2.5)
        assert ok == (prim[i] >= 0)
        if ok:
            assert h["primIndex"] == prim[i] and np.float32(h["t"]).view(np.uint32) == t[i].view(np.uint32)
        else:
            assert h["primIndex"] == -1
    gc.destroy()


def test_legacy_seam_instances_and_index_sanitising(pkg, ob):
    rng = np.random.default_rng(9)
    v0, t0 = _soup(rng, 300, extent=0.8, size=0.2)
    v1, t1 = _soup(rng, 200, extent=0.5, size=0.3)
    t0 = t0.copy(); t0[5] = (10_000, 1, 2); t0[7] = (-3, 4, 5)         # embree_connect.cpp:98-109 -> (vertexNum-1) x3
    gc = pkg.GeometryCore(); gc.init_and_clear()
    assert gc.add_mesh_3f(v0, t0) == 0 and gc.add_mesh_3f(v1, t1) == 1
    mats = []
    for k in range(3):
        M = np.eye(4, dtype=np.float32); M[:3, 3] = rng.uniform(-1.5, 1.5, 3); M[0, 0] = 1.0 + 0.25 * k; M[1, 2] = 0.1 * k
        mats.append(M)
    gc.instance_meshes(0, np.stack(mats[:2]).reshape(-1, 16))
    gc.instance_meshes(1, mats[2].reshape(1, 16))
    gc.commit_scene()
    # flattened world-space copy, same binary32 operation order as the commit (row-major 3x4)
    ts0 = t0.copy(); ts0[5] = len(v0) - 1; ts0[7] = len(v0) - 1
    fv, ft, fprim, finst = [], [], [], []
    base = 0
    for ii, (vv, tt, M) in enumerate([(v0, ts0, mats[0]), (v0, ts0, mats[1]), (v1, t1, mats[2])]):
        x, y, z = vv[:, 0], vv[:, 1], vv[:, 2]
        w = np.stack([M[r, 0] * x + M[r, 1] * y + M[r, 2] * z + M[r, 3] for r in range(3)], 1).astype(np.float32)
        fv.append(w); ft.append(tt + base); fprim.append(np.arange(len(tt))); finst.append(np.full(len(tt), ii)); base += len(vv)
    fv = np.ascontiguousarray(np.concatenate(fv)); ft = np.ascontiguousarray(np.concatenate(ft).astype(np.int32))
    fprim = np.concatenate(fprim); finst = np.concatenate(finst)
    rays = _rays(rng, 20000, pkg)
    found, hits = gc.closest_hit_batch(rays, 0.0, 1000.0)
    prim, t, u, v = _brute(ob, fv, ft, rays, 0.0, 1000.0)
    m = prim >= 0
    assert np.array_equal(found != 0, m) and m.sum() > 200
    assert np.array_equal(hits["primIndex"][m], fprim[prim[m]])
    assert np.array_equal(hits["instIndex"][m], finst[prim[m]])
    assert np.all(hits["geomIndex"][m] == 0)
    assert np.array_equal(_bits(hits["t"][m]), _bits(t[m]))
    gc.destroy()


def test_lbvh_large_mesh_vs_cpu_bvh(pkg, ob):
    """C5-shaped mesh at reduced tessellation (196 608 triangles): 200 k random + primary-like rays vs the oracle's own CPU
    BVH (checked against brute force on a sample)"""
    syn = _syn(pkg)
    pos, nrm, tris = syn.cube_sphere(128)
    pos = np.ascontiguousarray(pos); tris = np.ascontiguousarray(tris)
    gc = pkg.GeometryCore(); gc.init_and_clear(); gc.add_mesh_3f(pos, tris); gc.commit_scene()
    rng = np.random.default_rng(3)
    n = 200_000
    rays = np.zeros(n, pkg.ray_dt)
    o = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 12.0], (n, 3))
    target = np.array([0.0, 1.6, 2.5]) + rng.normal(size=(n, 3)) * 0.9
    d = target - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"], rays["direction"] = o.astype(np.float32), d.astype(np.float32)
    found, hits = gc.closest_hit_batch(rays, 0.0, 1000000.0)
    L = ob.lib()
    bvh = L.ao_bvh_build(pos.ctypes.data, len(pos), tris.ctypes.data, len(tris))
    prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    rr = np.ascontiguousarray(rays.view(ob.ray_dt))
    L.ao_bvh_closest(bvh, n, rr.ctypes.data, 0.0, 1000000.0, prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data, 0)
    L.ao_bvh_free(bvh)
    _check(found, hits, prim, t, u, v)
    assert 0.3 < (prim >= 0).mean() < 0.95
    sub = slice(0, 300)                                                 # the CPU BVH itself against brute force
    p2, t2, u2, v2 = _brute(ob, pos, tris, rays[sub], 0.0, 1000000.0)
    assert np.array_equal(p2, prim[sub]) and np.array_equal(_bits(t2), _bits(t[sub]))
    gc.destroy()


def _mesh_scene(pkg, ob, nseg, mode):
    """Cornell scene whose mesh is a tessellated, displaced sphere (nearest-hit semantics), set on both sides"""
    syn = _syn(pkg)
    pos, nrm, tris = syn.cube_sphere(nseg, radius=0.9, center=(0.3, 1.2, 2.6))
    osc = ob.OracleScene()
    osc.set_mesh(pos, nrm, tris)
    pkg.set_partition(0, 1)
    gs = pkg.Scene(); gs.Init()
    gs.set_mesh(pos, nrm, tris, mesh_mode=mode)
    return gs, osc


@pytest.mark.parametrize("nseg", [2, 24])              # 48 triangles: constant-memory loop; 6912 triangles: LBVH
def test_render_with_nearest_mesh_bit_exact(pkg, ob, nseg):
    gs, osc = _mesh_scene(pkg, ob, nseg, pkg.MESH_NEAREST)
    w, h = 96, 72
    g = gs.primary_hits(w, h); o = osc.primary_hits(w, h, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), f
    assert ((g["prim_type"] == 2) & (g["is_hit"] == 1)).sum() > 100
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 31
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = osc.render(ob.make_params(w, h, ob.PT_MIS, passes=2, seed=31, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert int(rt.last_stats.rays) == rays


def test_reference_bf_mode_rejects_big_meshes(pkg, ob):
    syn = _syn(pkg)
    pos, nrm, tris = syn.cube_sphere(8)
    gs = pkg.Scene(); gs.Init()
    with pytest.raises(pkg.GCoreError):
        gs.set_mesh(pos, nrm, tris, mesh_mode=pkg.MESH_REFERENCE_BF)    # brute-force reference semantics: <= 64 triangles
    gs.Init()


@pytest.mark.parametrize("seed,nsph", [(5, 96), (7, 16)])
def test_many_spheres_with_lbvh_mesh(pkg, ob, seed, nsph):
    """IntersectAllSpheres (geometry.adb:48-115) at the library's limit of 96 spheres next to an LBVH mesh: spheres of very different
    sizes, overlapping, partly outside the Cornell box; rays from inside and outside their bounds, axis-parallel, lying in and starting
    on the walls of an 8^3 grid over them (the stress set of the sphere-grid experiment, profiles/tools/experiments/
    r02_sphere_grid_and_prefetch.patch -- correct, but slower than the plain loop) -- every Hit field against the oracle, then a small
    PT_MIS render."""
    gs, osc = _mesh_scene(pkg, ob, 24, pkg.MESH_NEAREST)
    rng = np.random.default_rng(seed)
    sp = np.zeros(nsph, pkg.sphere_dt)
    sp["pos"] = rng.uniform([-3.2, -0.6, -0.8], [3.2, 5.6, 6.0], (nsph, 3)).astype(np.float32)
    sp["r"] = np.where(rng.random(nsph) < 0.2, rng.uniform(0.4, 1.1, nsph), rng.uniform(0.02, 0.25, nsph)).astype(np.float32)
    sp["mat"] = rng.choice([1, 5, 8, 0, 3], nsph)
    sp[nsph - 1] = ((0.0, 4.5, 1.0), 0.5, 4)                                     # the sphere light stays the last sphere
    d0 = osc.describe()
    gs.set_analytic(sp, d0["cornell"]); osc.set_spheres(sp.view(ob.sphere_dt))
    # an 8^3 grid over the spheres' padded bounds, to aim rays at its walls
    lo = (sp["pos"].astype(np.float64) - sp["r"][:, None]).min(0); hi = (sp["pos"].astype(np.float64) + sp["r"][:, None]).max(0)
    pad = 1e-2 * (hi - lo).max() + 1e-3
    lo -= pad; hi += pad
    cell = (hi - lo) / 8
    n = 400_000
    rr = np.zeros(n, ob.ray_dt)
    o = rng.uniform(lo - 1.0, hi + 1.0, (n, 3))
    tgt = rng.uniform(lo, hi, (n, 3))
    d = tgt - o
    k = rng.integers(0, 6, n)
    walls = lo + np.round((o - lo) / cell) * cell                                # nearest cell wall / edge / corner
    on1 = k == 1; o[on1, 0] = walls[on1, 0]                                      # origin on an x wall
    on2 = k == 2; o[on2] = walls[on2]                                            # origin on a cell corner
    ax = k == 3; a = rng.integers(0, 3, n); d[ax] = 0.0; d[ax, a[ax]] = rng.choice([-1.0, 1.0], int(ax.sum()))      # axis-parallel
    inw = k == 4; o[inw, 1] = walls[inw, 1]; d[inw, 1] = 0.0                     # lying in a y wall
    dg = k == 5; d[dg] = np.sign(d[dg]) * cell                                   # along the cells' diagonals: ties in the DDA
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rr["origin"] = o.astype(np.float32); rr["direction"] = d.astype(np.float32)
    g = gs.Find_Closest_Hit(rr.view(pkg.ray_dt))
    oh = osc.closest_hit_batch(rr, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a_, b_ = g[f], oh[f]
        if a_.dtype.kind == "f":
            a_, b_ = _bits(a_), _bits(b_)
        assert np.array_equal(a_, b_), f
    assert ((g["prim_type"] == 1) & (g["is_hit"] == 1)).sum() > 20_000
    w, h = 128, 96
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 77
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = osc.render(ob.make_params(w, h, ob.PT_MIS, passes=2, seed=77, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert int(rt.last_stats.rays) == rays
    gs.Init()
