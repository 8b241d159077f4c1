"""The scene that arrives through the reference's own seam (gcore_add_mesh_3f / gcore_instance_meshes / gcore_commit_scene,
cpp/embree_connect.cpp:69-184, 241-244, called by scene_hydra_embree.adb:253-296, 380-386) is traced AND shaded by the render
path: several meshes, instances, one material per mesh, optional vertex normals -- SURVEY 8(f3).  The oracle gets the same
scene flattened to world space (unshared vertices, per-triangle materials); ids, t and every accumulator bit must agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy(); b[np.isnan(a)] = 0x7FC00000
    return b


def _box_mesh(lo, hi):
    """closed box, CCW-outward triangles (12)"""
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    v = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]], np.float32)
    t = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [3, 7, 6], [3, 6, 2], [0, 4, 7], [0, 7, 3], [1, 2, 6], [1, 6, 5]], np.int32)
    return v, t


def _ico_sphere(n_lat=10, n_lon=16, r=0.5):
    """lat/long sphere with smooth vertex normals, CCW-outward"""
    vs, ns = [], []
    for i in range(n_lat + 1):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            vs.append(r * n); ns.append(n)
    tris = []
    for i in range(n_lat):
        for j in range(n_lon):
            a = i * n_lon + j; b = i * n_lon + (j + 1) % n_lon; c = (i + 1) * n_lon + j; d = (i + 1) * n_lon + (j + 1) % n_lon
            tris.append((a, b, d)); tris.append((a, d, c))
    return np.array(vs, np.float32), np.array(ns, np.float32), np.array(tris, np.int32)


def _matrix(tx, ty, tz, s=1.0, rot_y=0.0):
    M = np.eye(4, dtype=np.float32)
    c, sn = np.float32(np.cos(rot_y)), np.float32(np.sin(rot_y))
    M[0, 0] = c * s; M[0, 2] = sn * s; M[1, 1] = s; M[2, 0] = -sn * s; M[2, 2] = c * s
    M[:3, 3] = (tx, ty, tz)
    return M


@pytest.fixture()
def committed(pkg, ob):
    """two meshes, four instances, three materials: committed through the legacy seam and mirrored into the oracle"""
    pkg.set_partition(0, 1)
    gs = pkg.Scene(); gs.Init()                                       # materials, light, camera, box, spheres of scene.adb
    gs.set_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))   # no gcore_set_mesh mesh
    gc = pkg.GeometryCore(); gc.init_and_clear()
    bv, bt = _box_mesh((-0.4, 0.0, -0.4), (0.4, 0.8, 0.4))
    sv, sn, st = _ico_sphere()
    assert gc.add_mesh_3f(bv, bt) == 0 and gc.add_mesh_3f(sv, st) == 1
    gc.set_mesh_material(0, 3)                                        # red Lambert boxes, geometric normals
    gc.set_mesh_material(1, 8)                                        # Phong spheres, smooth vertex normals
    gc.set_mesh_normals_3f(1, sn)
    mats = [_matrix(-0.2, 0.0, 3.6, 1.0, 0.4), _matrix(1.0, 0.0, 1.2, 1.5, -0.3)]
    gc.instance_meshes(0, np.stack(mats).reshape(-1, 16))
    gc.instance_meshes(1, np.stack([_matrix(-0.2, 1.3, 3.6, 1.0), _matrix(0.3, 2.6, 2.0, 0.8, 1.0)]).reshape(-1, 16))
    gc.commit_scene()
    gc.use_committed_scene(True)
    pos, tris, n9, tmat = gc.committed_scene()
    assert len(tris) == 2 * 12 + 2 * len(st) and set(np.unique(tmat)) == {3, 8}
    # oracle: the same flattened scene with unshared vertices (per-corner normals), nearest-hit semantics, per-triangle materials
    osc = ob.OracleScene()
    upos = np.ascontiguousarray(pos[tris.reshape(-1)])
    unrm = np.ascontiguousarray(n9.reshape(-1, 3))
    utris = np.arange(3 * len(tris), dtype=np.int32).reshape(-1, 3)
    osc.set_mesh(upos, unrm, utris)
    osc.set_mesh_tri_materials(tmat)
    yield gs, gc, osc, (pos, tris, n9, tmat)
    gc.use_committed_scene(False)
    gc.destroy()
    gs.Init()


def test_flattening_matches_numpy(committed):
    """world-space vertices = 3x4 row-major instance matrix applied in the commit's operation order; geometric normals of the
    boxes = normalize(cross(B - A, C - A)); sphere normals = normalised 3x3 * vertex normal"""
    gs, gc, osc, (pos, tris, n9, tmat) = committed
    bv, bt = _box_mesh((-0.4, 0.0, -0.4), (0.4, 0.8, 0.4))
    M = _matrix(-0.2, 0.0, 3.6, 1.0, 0.4)
    x, y, z = bv[:, 0], bv[:, 1], bv[:, 2]
    w = np.stack([M[r, 0] * x + M[r, 1] * y + M[r, 2] * z + M[r, 3] for r in range(3)], 1).astype(np.float32)
    assert np.array_equal(_bits(pos[:8]), _bits(w))
    A, B, C = w[bt[:, 0]], w[bt[:, 1]], w[bt[:, 2]]
    e1, e2 = (B - A).astype(np.float32), (C - A).astype(np.float32)
    cr = np.stack([e1[:, 1] * e2[:, 2] - e2[:, 1] * e1[:, 2], e1[:, 2] * e2[:, 0] - e2[:, 2] * e1[:, 0], e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]], 1).astype(np.float32)
    d = (cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1] + cr[:, 2] * cr[:, 2]).astype(np.float32)
    n = (np.float32(1.0) / np.sqrt(d, dtype=np.float32))[:, None] * cr
    assert np.array_equal(_bits(n9[:12, 0]), _bits(n.astype(np.float32)))
    assert np.all(np.abs(np.linalg.norm(n9.reshape(-1, 3), axis=1) - 1.0) < 1e-5)


def test_committed_scene_queries_bit_exact(committed, pkg, ob):
    gs, gc, osc, (pos, tris, n9, tmat) = committed
    w, h = 160, 120
    g = gs.primary_hits(w, h); o = osc.primary_hits(w, h, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), f
    tri = (g["prim_type"] == 2) & (g["is_hit"] == 1)
    assert tri.sum() > 300 and set(np.unique(g["matId"][tri])) == {3, 8}     # both meshes visible, each with its own material
    rng = np.random.default_rng(11)
    n = 100_000
    rays = np.zeros(n, pkg.ray_dt)
    rays["origin"] = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 9.0], (n, 3)).astype(np.float32)
    d = np.array([0.2, 1.2, 2.6]) + rng.normal(size=(n, 3)) * 1.0 - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = gs.Find_Closest_Hit(rays); o = osc.closest_hit_batch(rays.view(ob.ray_dt), mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), f


@pytest.mark.parametrize("rtype,depth", [("PT_MIS", 8), ("PT_SHADOW", 4), ("PT_STUPID", 6), ("RT_DEBUG", 1)])
def test_committed_scene_renders_bit_exact(committed, pkg, ob, rtype, depth):
    """the integrators shade what the seam committed: materials 3 (Lambert) and 8 (Phong) on the meshes, everything else of
    scene.adb around them"""
    gs, gc, osc, _ = committed
    w, h = 112, 84
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.Max_Trace_Depth = depth; rt.seed = 808
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = osc.render(ob.make_params(w, h, getattr(ob, rtype), max_depth=depth, passes=2, seed=808, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    if rtype != "RT_DEBUG":
        assert int(rt.last_stats.rays) == rays
        assert np.array_equal(rt.screen_buffer, ob.resolve(acc, rt.GetSPP()))


def test_seam_still_answers_in_mesh_and_instance_ids(committed, pkg, ob):
    """gcore_closest_hit[_batch] on the same committed scene: primIndex inside the mesh, instIndex in declaration order"""
    gs, gc, osc, (pos, tris, n9, tmat) = committed
    rng = np.random.default_rng(5)
    n = 20_000
    rays = np.zeros(n, pkg.ray_dt)
    rays["origin"] = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 9.0], (n, 3)).astype(np.float32)
    d = np.array([0.2, 1.2, 2.6]) + rng.normal(size=(n, 3)) * 1.0 - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    found, hits = gc.closest_hit_batch(rays, 0.0, 1000.0)
    prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
    rr = np.ascontiguousarray(rays.view(ob.ray_dt))
    ob.lib().ao_bruteforce_closest(pos.ctypes.data, tris.ctypes.data, len(tris), n, rr.ctypes.data, 0.0, 1000.0, prim.ctypes.data, t.ctypes.data, None, None)
    m = prim >= 0
    assert np.array_equal(found != 0, m) and m.sum() > 1000
    nb, ns = 12, (len(tris) - 24) // 2
    inst = np.where(prim < 12, 0, np.where(prim < 24, 1, np.where(prim < 24 + ns, 2, 3)))
    local = np.where(prim < 24, prim % 12, (prim - 24) % ns)
    assert np.array_equal(hits["instIndex"][m], inst[m]) and np.array_equal(hits["primIndex"][m], local[m])
    assert np.array_equal(_bits(hits["t"][m]), _bits(t[m]))


def test_two_sided_seam_and_occlusion_query(pkg, ob):
    """gcore_set_seam_two_sided(1): Embree's rtcIntersect1 hits both faces; gcore_bvh_occluded_batch: rtcOccluded1 counterpart"""
    rng = np.random.default_rng(21)
    c = rng.uniform(-2, 2, (1500, 1, 3))
    verts = (c + rng.uniform(-0.3, 0.3, (1500, 3, 3))).astype(np.float32).reshape(-1, 3)
    tris = np.arange(4500, dtype=np.int32).reshape(-1, 3)
    gc = pkg.GeometryCore(); gc.init_and_clear(); gc.add_mesh_3f(verts, tris); gc.commit_scene()
    n = 30_000
    rays = np.zeros(n, pkg.ray_dt)
    rays["origin"] = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rr = np.ascontiguousarray(rays.view(ob.ray_dt))

    def brute(fn, t0, t1):
        prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        fn(verts.ctypes.data, tris.ctypes.data, len(tris), n, rr.ctypes.data, t0, t1, prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        return prim, t, u, v

    p1, t1, _, _ = brute(ob.lib().ao_bruteforce_closest, 0.0, 100000.0)
    p2, t2, u2, v2 = brute(ob.lib().ao_bruteforce_closest_two_sided, 0.0, 100000.0)
    assert (p2 >= 0).sum() > 1.2 * (p1 >= 0).sum()                      # back faces are hit as well
    gc.set_two_sided(True)
    found, hits = gc.closest_hit_batch(rays, 0.0, 100000.0)
    m = p2 >= 0
    assert np.array_equal(found != 0, m)
    # exact t-ties between different triangles are resolved by the lowest index on both sides; compare ids where t is unique
    assert np.array_equal(hits["primIndex"][m], p2[m]) and np.array_equal(_bits(hits["t"][m]), _bits(t2[m]))
    assert np.array_equal(_bits(hits["texCoord"][m, 0]), _bits(v2[m])) and np.array_equal(_bits(hits["texCoord"][m, 1]), _bits(u2[m]))
    occ = gc.occluded_batch(rays, 0.0, 100000.0)
    assert np.array_equal(occ != 0, m)
    gc.set_two_sided(False)
    found, hits = gc.closest_hit_batch(rays, 0.0, 100000.0)
    assert np.array_equal(found != 0, p1 >= 0) and np.array_equal(hits["primIndex"][p1 >= 0], p1[p1 >= 0])
    occ = gc.occluded_batch(rays, 0.5, 2.0)                              # a window: any hit inside (0.5, 2.0)
    pw, _, _, _ = brute(ob.lib().ao_bruteforce_closest, 0.5, 2.0)
    assert np.array_equal(occ != 0, pw >= 0)
    gc.destroy()
