"""BASELINE config C5: Cornell box + sphere light + 64 analytic spheres + a 10 002 828-triangle displaced
cube-sphere through the GPU LBVH.  Acceptance: triangle ids and t/u/v of random + primary rays equal the oracle's CPU BVH
exactly; a small render of the full scene equals the oracle's bit for bit."""
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


@pytest.fixture(scope="module")
def c5(pkg, ob):
    syn = _syn(pkg)
    pkg.set_partition(0, 1)
    gs = pkg.Scene(); gs.Init()
    osc = ob.OracleScene()
    pos, nrm, tris = syn.install_c5(gs, 913, osc)
    assert len(tris) == 10_002_828
    yield gs, osc, np.ascontiguousarray(pos), np.ascontiguousarray(tris)
    gs.Init()                                                         # back to the pyramid2 scene for the other tests


def test_c5_scene_queries_match_oracle(c5, pkg, ob):
    """SURVEY 8d C5 acceptance: 1 M random + 1 M primary rays against the oracle's CPU BVH, every field of the Hit record exact.
    Primary rays: the 1366x768 frame of the scene's camera (1 049 088 pixel-centre rays) through gcore_primary_hits; random rays:
    origins uniform in the box, aimed at a Gaussian cloud around the mesh, through gcore_closest_hit_batch."""
    gs, osc, pos, tris = c5
    w, h = 1366, 768
    g = gs.primary_hits(w, h); o = osc.primary_hits(w, h, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), "primary " + f
    tri = (g["prim_type"] == 2) & (g["is_hit"] == 1)
    assert tri.sum() > 30_000 and len(np.unique(g["prim_index"][tri])) > 20_000
    rng = np.random.default_rng(8)
    n = 1_000_000
    rr = np.zeros(n, ob.ray_dt)
    rr["origin"] = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 12.0], (n, 3)).astype(np.float32)
    target = np.array([0.0, 1.6, 2.5]) + rng.normal(size=(n, 3)) * 1.2
    d = target - rr["origin"]; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rr["direction"] = d.astype(np.float32)
    g = gs.Find_Closest_Hit(rr.view(pkg.ray_dt))
    o = osc.closest_hit_batch(rr, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), "random " + f
    tri = (g["prim_type"] == 2) & (g["is_hit"] == 1)
    assert tri.sum() > 150_000 and len(np.unique(g["prim_index"][tri])) > 60_000


def test_c5_primary_hits_and_small_render_bit_exact(c5, pkg, ob):
    gs, osc, pos, tris = c5
    w, h = 384, 216
    g = gs.primary_hits(w, h); o = osc.primary_hits(w, h, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), f
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 0x5EED0005
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = osc.render(ob.make_params(w, h, ob.PT_MIS, passes=2, seed=0x5EED0005, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert int(rt.last_stats.rays) == rays


def test_c5_bvh_facts(c5, pkg):
    gc = pkg.GeometryCore()
    info = gc.bvh_info()
    assert info["ntris"] == 10_002_828 and info["nnodes"] == 10_002_827
    assert info["bytes"] == 10_002_828 * (48 + 36) + 10_002_827 * 32        # 48-B triangle slots, 36 B of corner normals, 32-B quantised nodes


def test_c5_axis_parallel_rays(c5, pkg, ob):
    """Rays with one or two exactly-zero direction components (1/d = +-inf; a cosine sample with r2 = 0 off an axis-aligned wall makes
    one).  The hit must equal the oracle's, and the walk must stay as short as that of the same ray tilted by 1e-4: an axis that drops out
    of the slab test (inf * 0 = NaN) is still correct but turns a line of boxes into a slab of the tree."""
    gs, osc, pos, tris = c5
    rng = np.random.default_rng(77)
    n = 24_000
    rr = np.zeros(n, ob.ray_dt)
    rr["origin"] = rng.uniform([-1.6, 0.2, 0.6], [1.6, 3.2, 4.4], (n, 3)).astype(np.float32)
    d = np.zeros((n, 3), np.float32)
    kind = rng.integers(0, 2, n)                                      # 0: two zero components, 1: one zero component
    ax = rng.integers(0, 3, n); sg = rng.choice(np.float32([-1.0, 1.0]), n)
    d[np.arange(n), ax] = sg
    ax2 = (ax + 1 + rng.integers(0, 2, n)) % 3
    sel = kind == 1
    d[np.arange(n)[sel], ax2[sel]] = rng.uniform(-1, 1, int(sel.sum())).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    neg0 = rng.random((n, 3)) < 0.3
    d[(d == 0) & neg0] = np.float32(-0.0)                             # +0.0 and -0.0 (1/d = +inf, -inf)
    rr["direction"] = d
    g = gs.Find_Closest_Hit(rr.view(pkg.ray_dt))
    o = osc.closest_hit_batch(rr, mesh_mode=ob.MESH_NEAREST)
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), "axis-parallel " + f
    assert ((g["prim_type"] == 2) & (g["is_hit"] == 1)).sum() > 1000
    tilted = rr.copy()
    t = d + np.float32(1.0e-4) * rng.normal(size=(n, 3)).astype(np.float32)
    tilted["direction"] = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(np.float32)
    a = pkg.bvh_bench(rr.view(pkg.ray_dt), repeat=1); b = pkg.bvh_bench(tilted.view(pkg.ray_dt), repeat=1)
    assert a["nodes_per_ray"] < 1.25 * b["nodes_per_ray"] + 4.0, (a, b)
