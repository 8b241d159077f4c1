"""GPU parity tests for the kernel instantiations and configurations that round 1 left unexercised:
LBVH render kernels for PT_SHADOW / PT_STUPID / RT_DEBUG, brute-force reference meshes of 9..64 triangles (the non-pair
mesh_query loop), tile-partition independence on an LBVH scene, a pass that follows a debug pass without a resize, and the
re-entrancy of the legacy one-ray seam under 16 host threads."""
import ctypes as C
import importlib.util
import os
import threading

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


def _mesh_scene(pkg, ob, nseg, mode, radius=0.9, center=(0.3, 1.2, 2.6)):
    syn = _syn(pkg)
    pos, nrm, tris = syn.cube_sphere(nseg, radius=radius, center=center)
    osc = ob.OracleScene()
    osc.set_mesh(pos, nrm, tris)
    pkg.set_partition(0, 1)
    gs = pkg.Scene(); gs.Init()
    gs.set_mesh(pos, nrm, tris, mesh_mode=mode)
    return gs, osc


def _hits_equal(g, o):
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        assert np.array_equal(a, b), f


@pytest.mark.parametrize("rtype,depth,units", [("PT_SHADOW", 8, 2), ("PT_SHADOW", 1, 3), ("PT_STUPID", 8, 2), ("PT_MIS", 5, 2)])
@pytest.mark.parametrize("aa", [1, 0])
def test_lbvh_scene_all_integrators_bit_exact(pkg, ob, rtype, depth, units, aa):
    """render_wave_kernel<PT_SHADOW|PT_STUPID|PT_MIS, LBVH>: a 6912-triangle mesh (through the GPU LBVH) in the Cornell scene"""
    gs, osc = _mesh_scene(pkg, ob, 24, pkg.MESH_NEAREST)
    w, h = 88, 66
    rt = pkg.Ray_Tracer(); rt.Threads_Num = units; rt.Max_Trace_Depth = depth; rt.Anti_Aliasing_On = bool(aa); rt.seed = 1234
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = osc.render(ob.make_params(w, h, getattr(ob, rtype), aa=aa, max_depth=depth, passes=units, seed=1234, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert int(rt.last_stats.rays) == rays
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, rt.GetSPP()))
    gs.Init()


def test_lbvh_scene_debug_palette(pkg, ob):
    """debug_kernel<LBVH> (ray_tracer.adb:208-238): palette(matId rem 8) of the primary hit, raw gamma without 1/spp"""
    gs, osc = _mesh_scene(pkg, ob, 24, pkg.MESH_NEAREST)
    w, h = 160, 120
    rt = pkg.Ray_Tracer()
    rt.Init_Render(pkg.RT_DEBUG); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, _ = osc.render(ob.make_params(w, h, ob.RT_DEBUG, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, 1))
    assert (acc[..., 2] == 0.5).sum() > 500                         # matId 2 = the mesh is visible
    gs.Init()


@pytest.mark.parametrize("nseg,ntris", [(1, 12), (2, 48)])
@pytest.mark.parametrize("mode", ["REFERENCE_BF", "NEAREST"])
def test_const_mesh_9_to_64_triangles(pkg, ob, nseg, ntris, mode):
    """constant-memory meshes above the 8-triangle pair stage: the one-ray-per-thread mesh_query loop, in both window modes
    (IntersectMeshBF's (t, t + 1e-6) window, geometry.adb:289-293, and nearest hit)"""
    gmode = getattr(pkg, "MESH_" + mode); omode = getattr(ob, "MESH_" + mode)
    gs, osc = _mesh_scene(pkg, ob, nseg, gmode, radius=1.0, center=(0.2, 1.4, 2.8))
    assert len(gs.get_mesh()[2]) == ntris
    w, h = 96, 72
    _hits_equal(gs.primary_hits(w, h), osc.primary_hits(w, h, mesh_mode=omode))
    rng = np.random.default_rng(ntris)
    rays = np.zeros(40_000, pkg.ray_dt)
    rays["origin"] = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 8.0], (len(rays), 3)).astype(np.float32)
    d = np.array([0.2, 1.4, 2.8]) + rng.normal(size=(len(rays), 3)) * 0.8 - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = gs.Find_Closest_Hit(rays); o = osc.closest_hit_batch(rays.view(ob.ray_dt), mesh_mode=omode)
    _hits_equal(g, o)
    assert ((g["prim_type"] == 2) & (g["is_hit"] == 1)).sum() > 2000
    for rtype in ("PT_MIS", "PT_SHADOW", "PT_STUPID"):
        rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 4321
        rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
        hdr = rt.Render_Pass(read_hdr=True)
        acc, nr = osc.render(ob.make_params(w, h, getattr(ob, rtype), passes=2, seed=4321, mesh_mode=omode))
        assert np.array_equal(_bits(hdr[..., :3]), _bits(acc)), rtype
        assert int(rt.last_stats.rays) == nr
    gs.Init()


def _gather_partition(pkg, rt, w, h, units, world, spp):
    """render `units` units with every rank of a `world`-way partition in turn and resolve the gathered accumulators"""
    import torch
    bufs = []
    for rank in range(world):
        pkg.set_partition(rank, world)
        rt.Resize_Viewport(w, h)
        cfg = rt._cfg()
        assert pkg.lib().gcore_init_render(C.byref(cfg)) == 0
        st = pkg.RenderStats()
        assert pkg.lib().gcore_render_units(units, C.byref(st)) == 0
        ptr, nslots = pkg.accum_device_ptr()
        t = torch.empty(nslots * 4, dtype=torch.float32, device="cuda")
        pkg.synchronize()
        C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr), C.c_size_t(nslots * 16), 3)
        bufs.append(t)
    allb = torch.cat(bufs)
    hdr = torch.empty(h * w * 4, dtype=torch.float32, device="cuda")
    ldr = torch.empty(h * w, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert pkg.lib().gcore_resolve_gathered_device(C.c_void_p(allb.data_ptr()), spp, C.c_void_p(hdr.data_ptr()), C.c_void_p(ldr.data_ptr())) == 0
    pkg.synchronize()
    pkg.set_partition(0, 1)
    return hdr.cpu().numpy().reshape(h, w, 4), ldr.cpu().numpy().view(np.uint32).reshape(h, w)


def test_partition_independence_lbvh_scene(pkg, ob):
    """SURVEY 8e on an LBVH scene: worlds 2, 3, 8 give the bits of world 1 (and of the oracle)"""
    gs, osc = _mesh_scene(pkg, ob, 24, pkg.MESH_NEAREST)
    w, h = 168, 104
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 555
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    ref = rt.Render_Pass(read_hdr=True).copy(); ref_ldr = rt.screen_buffer.copy()
    acc, _ = osc.render(ob.make_params(w, h, ob.PT_MIS, passes=2, seed=555, mesh_mode=ob.MESH_NEAREST))
    assert np.array_equal(_bits(ref[..., :3]), _bits(acc))
    for world in (2, 3, 8):
        got, ldr = _gather_partition(pkg, rt, w, h, 2, world, 8)
        assert np.array_equal(_bits(got[..., :3]), _bits(ref[..., :3])), "world=%d" % world
        assert np.array_equal(ldr, ref_ldr)
    gs.Init()


def test_debug_pass_after_pt_pass_without_resize(gscene, oscene, pkg, ob):
    """Init_Render(PT_*), Render_Pass, Init_Render(RT_DEBUG), Render_Pass.  The reference overwrites g_accBuff with
    the palette, resolves it WITHOUT 1/spp and never touches g_spp (ray_tracer.adb:245-260); a later PT pass then adds onto
    the palette and divides by the untouched spp."""
    w, h = 96, 64
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 3; rt.seed = 2024
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    rt.Render_Pass()
    assert rt.GetSPP() == 12
    rt.Init_Render(pkg.RT_DEBUG)
    hdr = rt.Render_Pass(read_hdr=True)
    assert rt.GetSPP() == 12                                          # g_spp untouched
    pal, _ = oscene.render(ob.make_params(w, h, ob.RT_DEBUG))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(pal))
    assert np.array_equal(rt.screen_buffer, ob.resolve(pal, 1))       # palette ** 0.5, no normC
    # and on a fresh viewport a debug pass must not move spp off 0: the next PT pass starts at sample 0 and divides by 4N
    rt.Resize_Viewport(w, h)
    rt.Render_Pass()
    assert rt.GetSPP() == 0
    rt.Init_Render(pkg.PT_MIS); rt.g_finish = False
    hdr = rt.Render_Pass(read_hdr=True)
    assert rt.GetSPP() == 12
    acc, _ = oscene.render(ob.make_params(w, h, ob.PT_MIS, passes=3, seed=2024), accum=pal.copy())
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, 12))


def test_closest_hit_seam_16_threads(pkg, ob):
    """scene_hydra_embree.adb:426-446: Find_Closest_Hit -> gcore_closest_hit is called concurrently from Threads_Num Ada tasks
    with no locking on the caller's side.  16 host threads hammer the one-ray entry point; every answer equals brute force."""
    rng = np.random.default_rng(77)
    c = rng.uniform(-2, 2, (800, 1, 3))
    verts = (c + rng.uniform(-0.3, 0.3, (800, 3, 3))).astype(np.float32).reshape(-1, 3)
    tris = np.arange(2400, dtype=np.int32).reshape(-1, 3)
    gc = pkg.GeometryCore(); gc.init_and_clear(); gc.add_mesh_3f(verts, tris); gc.commit_scene()
    nthreads, per = 16, 250
    n = nthreads * per
    rays = np.zeros(n, pkg.ray_dt)
    rays["origin"] = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    rr = np.ascontiguousarray(rays.view(ob.ray_dt))
    ob.lib().ao_bruteforce_closest(verts.ctypes.data, tris.ctypes.data, len(tris), n, rr.ctypes.data, 0.0, 100000.0,
                                   prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
    got_prim = np.full(n, -7, np.int32); got_t = np.zeros(n, np.float32); got_ok = np.zeros(n, np.int32)
    L = pkg.lib()
    errors = []

    def work(k):
        try:
            h = np.zeros(1, pkg.hitcpp_dt)
            for i in range(k * per, (k + 1) * per):
                h["primIndex"] = -1; h["t"] = 0
                o = np.ascontiguousarray(rays["origin"][i]); dd = np.ascontiguousarray(rays["direction"][i])
                ok = L.gcore_closest_hit(o.ctypes.data, dd.ctypes.data, 0.0, 100000.0, h.ctypes.data)
                got_ok[i] = 1 if ok else 0; got_prim[i] = h["primIndex"][0]; got_t[i] = h["t"][0]
        except Exception as e:                                         # pragma: no cover
            errors.append(e)

    sv_before = np.zeros(2, np.uint64); st_before = np.zeros(2, np.uint64)
    L.gcore_seam_server_stats(sv_before.ctypes.data); L.gcore_seam_stats(st_before.ctypes.data)
    th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors
    m = prim >= 0
    assert m.sum() > 100
    assert np.array_equal(got_ok != 0, m)
    assert np.array_equal(got_prim[m], prim[m]) and np.all(got_prim[~m] == -1)
    assert np.array_equal(_bits(got_t[m]), _bits(t[m]))
    sv = np.zeros(2, np.uint64); st = np.zeros(2, np.uint64)
    assert L.gcore_seam_server_stats(sv.ctypes.data) == 0 and L.gcore_seam_stats(st.ctypes.data) == 0
    if os.environ.get("GCORE_SEAM_SERVER", "1") != "0":
        # every batch went through the resident kernel, not the launch fall-back
        assert sv[0] >= 1 and sv[1] - sv_before[1] == st[0] - st_before[0] and st[1] - st_before[1] == n, (sv_before, sv, st_before, st)
    gc.destroy()


def test_seam_resident_kernel_sees_scene_changes(pkg, ob):
    """The resident seam kernel reads the scene from constant memory: it must be gone before that block changes.  One ray at a time
    (the kernel stays resident in between), then a different scene is committed, then the two-sided test is switched on -- each
    time every answer equals brute force on the scene of the moment; a pause longer than the kernel's idle limit in between."""
    import time
    rng = np.random.default_rng(5)
    L = pkg.lib()
    gc = pkg.GeometryCore()

    def soup(nt, off):
        c = rng.uniform(-2, 2, (nt, 1, 3)) + off
        return (c + rng.uniform(-0.3, 0.3, (nt, 3, 3))).astype(np.float32).reshape(-1, 3), np.arange(3 * nt, dtype=np.int32).reshape(-1, 3)

    def ask(rays):
        h = np.zeros(1, pkg.hitcpp_dt); out = np.full(len(rays), -1, np.int32); tt = np.zeros(len(rays), np.float32)
        for i in range(len(rays)):
            o = np.ascontiguousarray(rays["origin"][i]); d = np.ascontiguousarray(rays["direction"][i])
            if L.gcore_closest_hit(o.ctypes.data, d.ctypes.data, 0.0, 100000.0, h.ctypes.data):
                out[i] = h["primIndex"][0]; tt[i] = h["t"][0]
        return out, tt

    def brute(verts, tris, rays, two_sided):
        n = len(rays); prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        rr = np.ascontiguousarray(rays.view(ob.ray_dt))
        f = ob.lib().ao_bruteforce_closest_two_sided if two_sided else ob.lib().ao_bruteforce_closest
        f(verts.ctypes.data, tris.ctypes.data, len(tris), n, rr.ctypes.data, 0.0, 100000.0, prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        return prim, t

    n = 600
    rays = np.zeros(n, pkg.ray_dt)
    rays["origin"] = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    sv0 = np.zeros(2, np.uint64); L.gcore_seam_server_stats(sv0.ctypes.data)
    for k, (nt, off, two) in enumerate([(1500, 0.0, False), (2500, 0.4, False), (2500, 0.4, True), (1200, -0.3, False)]):
        if k != 2:
            verts, tris = soup(nt, off)
            gc.init_and_clear(); gc.add_mesh_3f(verts, tris); gc.commit_scene()
        gc.set_two_sided(two)
        if k == 3:
            time.sleep(0.05)                                           # well past the 2 ms idle limit: the kernel has left and is started again
        got, gt = ask(rays)
        want, wt = brute(verts, tris, rays, two)
        m = want >= 0
        assert m.sum() > 40
        assert np.array_equal(got, np.where(m, want, -1)), k
        assert np.array_equal(_bits(gt[m]), _bits(wt[m])), k
    gc.set_two_sided(False)
    sv1 = np.zeros(2, np.uint64); L.gcore_seam_server_stats(sv1.ctypes.data)
    if os.environ.get("GCORE_SEAM_SERVER", "1") != "0":
        assert sv1[1] - sv0[1] >= 4 * n - 4 and sv1[0] - sv0[0] >= 4, (sv0, sv1)
    gc.destroy()
