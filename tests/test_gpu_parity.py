"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact everywhere (integer ids AND binary32 values): the kernels are compiled without FMA contraction and with IEEE
div/sqrt, and use the same Philox stream and deterministic sin/cos/pow as the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    """binary32 bit patterns; NaNs are canonicalised: the NaN *payload/sign* is not part of the contract (x86 SSE produces
    the default NaN 0xFFC00000, NVIDIA GPUs 0x7FFFFFFF) -- a NaN must still be a NaN on both sides."""
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def _assert_hits_equal(g, o):
    for f in g.dtype.names:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            a, b = _bits(a), _bits(b)
        bad = np.argwhere(a != b)
        assert bad.size == 0, "field %s differs at %d records, first %s: gpu %s oracle %s" % (
            f, len(bad), bad[0], g[tuple(bad[0][:g.ndim])], o[tuple(bad[0][:g.ndim])])


def test_scene_setup_matches_oracle(gscene, oscene):
    """world-space mesh as the device sees it == the oracle's LoadMeshFromVSGF + transform (scene.adb:194-206)"""
    pos, nrm, tris, bbox = gscene.get_mesh()
    d = oscene.describe()
    assert np.array_equal(_bits(pos), _bits(d["pos"]))
    assert np.array_equal(_bits(nrm), _bits(d["nrm"]))
    assert np.array_equal(tris, d["tris"])
    assert np.array_equal(_bits(bbox), _bits(d["bbox"]))


@pytest.mark.parametrize("w,h", [(1024, 1024), (1024, 768), (100, 37), (33, 65)])
def test_primary_hits_bit_exact(gscene, oscene, w, h):
    """BASELINE config C2 (1024x1024, pixel-centre rays): every field of the Hit record, every pixel."""
    g = gscene.primary_hits(w, h)
    o = oscene.primary_hits(w, h)
    _assert_hits_equal(g, o)
    if (w, h) == (1024, 1024):
        assert 0.43 < g["is_hit"].mean() < 0.46
        tri = g[(g["is_hit"] == 1) & (g["prim_type"] == 2)]
        assert set(np.unique(tri["prim_index"])) == {0, 1}          # only side faces 0 and 1 face the camera


def _random_rays(rng, n, pkg):
    rays = np.zeros(n, pkg.ray_dt)
    o = rng.uniform([-2.4, 0.1, 0.1], [2.4, 4.9, 12.0], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # a share of degenerate directions: axis-parallel rays make 1/0 = inf and 0*inf = NaN in the slab tests
    k = n // 8
    d[:k] = 0
    d[np.arange(k), rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)
    # rays starting inside the spheres and exactly on box faces
    o[k:2 * k] = np.array([1.4, 1.0, 3.0], np.float32) + rng.uniform(-0.5, 0.5, (k, 3)).astype(np.float32)
    o[2 * k:3 * k, 1] = 0.0
    o[3 * k:4 * k, 0] = -2.5
    rays["origin"], rays["direction"] = o, d.astype(np.float32)
    return rays


def test_closest_hit_batch_random_rays(gscene, oscene, pkg):
    rng = np.random.default_rng(1234)
    rays = _random_rays(rng, 200_000, pkg)
    g = gscene.Find_Closest_Hit(rays)
    o = oscene.closest_hit_batch(rays)
    _assert_hits_equal(g, o)
    assert 0.3 < g["is_hit"].mean() <= 1.0


def test_closest_hit_rect_light_scene(pkg, ob):
    """the AreaLight variant of scene.adb:109-114,130-136: IntersectFlatLight joins the candidates"""
    osc = ob.OracleScene()
    d = osc.describe()
    light = d["light"].copy(); light["shape"] = ob.AO_LIGHT_RECT if hasattr(ob, "AO_LIGHT_RECT") else 0
    light["intensity"] = 20.0
    light["surfaceArea"] = (0.75 + 0.75) * (3.25 - 1.25)
    fl = np.zeros(1, ob.flatlight_dt)
    fl["boxMin"] = light["boxMin"]; fl["boxMax"] = light["boxMax"]; fl["normal"] = light["normal"]; fl["intensity"] = 20.0
    osc.set_light(light); osc.set_flatlight(fl); osc.set_spheres(d["spheres"][:2])
    pkg.set_partition(0, 1)
    gs = pkg.Scene(); gs.Init()
    gs.set_light(light.view(pkg.light_dt)); gs.set_analytic(d["spheres"][:2].view(pkg.sphere_dt), d["cornell"].view(pkg.cornell_dt),
                                                              fl.view(pkg.flatlight_dt))
    rays = _random_rays(np.random.default_rng(7), 50_000, pkg)
    g = gs.Find_Closest_Hit(rays); o = osc.closest_hit_batch(rays)
    _assert_hits_equal(g, o)
    assert (g["prim_type"] == 3).sum() > 100                       # quads were actually hit
    # and a short render with the rect light, bit for bit
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 1
    for rtype in (pkg.PT_SHADOW, pkg.PT_MIS):
        rt.Init_Render(rtype); rt.Resize_Viewport(64, 48)
        hdr = rt.Render_Pass(read_hdr=True)
        acc, rays_o = osc.render(ob.make_params(64, 48, rtype, passes=1, seed=rt.seed))
        assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
        assert int(rt.last_stats.rays) == rays_o


RENDER_CASES = [
    # (render_type, aa, depth, units, w, h)
    ("PT_STUPID", 1, 8, 2, 80, 60),
    ("PT_SHADOW", 1, 1, 4, 96, 64),      # BASELINE config C3 shape (direct lighting only), reduced size
    ("PT_SHADOW", 1, 8, 2, 80, 60),
    ("PT_MIS", 1, 8, 3, 96, 64),         # BASELINE config C4 shape, reduced size
    ("PT_MIS", 0, 8, 5, 70, 50),         # Anti_Aliasing_On = false: pixel-centre rays, 1 spp per unit
    ("PT_MIS", 1, 3, 2, 64, 64),
    ("PT_MIS", 1, 16, 1, 48, 32),
]


@pytest.mark.parametrize("rtype,aa,depth,units,w,h", RENDER_CASES)
def test_render_bit_exact(gscene, oscene, pkg, ob, rtype, aa, depth, units, w, h):
    """integrator loop + NEE + shading + accumulation: HDR accumulator, ray count and LDR frame equal the oracle's"""
    rt = pkg.Ray_Tracer()
    rt.Threads_Num = units; rt.Anti_Aliasing_On = bool(aa); rt.Max_Trace_Depth = depth; rt.seed = 0x5EED0004
    rt.Background_Color = (0.0, 0.0, 0.0)
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    p = ob.make_params(w, h, getattr(ob, rtype), aa=aa, max_depth=depth, passes=units, seed=rt.seed)
    acc, rays = oscene.render(p)
    diff = (_bits(hdr[..., :3]) != _bits(acc)).any(-1)
    assert diff.sum() == 0, "%d of %d pixels differ; first at %s: gpu %s oracle %s" % (
        diff.sum(), w * h, np.argwhere(diff)[0], hdr[tuple(np.argwhere(diff)[0])], acc[tuple(np.argwhere(diff)[0])])
    assert int(rt.last_stats.rays) == rays
    assert rt.GetSPP() == units * (4 if aa else 1)
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, rt.GetSPP()))


@pytest.mark.parametrize("rtype,aa,depth,units", [("PT_MIS", 1, 8, 2), ("PT_SHADOW", 0, 4, 3), ("PT_STUPID", 1, 8, 1)])
def test_render_equals_the_numpy_restatement_directly(gscene, oscene, pkg, ob, rtype, aa, depth, units):
    """the CUDA path against tests/ada_numpy_integrators.py -- the restatement of the whole render that was written from the Ada
    sources independently of the oracle -- without the oracle in between (it only supplies the scene constants and sin / cos / **)"""
    import ada_numpy_integrators as ai
    w, h = 56, 40
    rt = pkg.Ray_Tracer()
    rt.Threads_Num = units; rt.Anti_Aliasing_On = bool(aa); rt.Max_Trace_Depth = depth; rt.seed = 424242
    rt.Background_Color = (0.0, 0.0, 0.0)
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    d = oscene.describe()
    sc = {"spheres": d["spheres"], "cornell": d["cornell"][0], "pos": d["pos"], "nrm": d["nrm"], "tris": d["tris"], "bbox": d["bbox"]}
    acc, rays = ai.do_pass(getattr(ob, rtype), w, h, units, bool(aa), depth, rt.seed, sc, d["materials"], d["light"][0],
                           {"pos": d["camera"]["pos"][0], "matrix": d["camera"]["matrix"][0]}, ob.sincos_batch, ob.expon_batch)
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert int(rt.last_stats.rays) == rays


def test_background_colour_and_progressive_passes(gscene, oscene, pkg, ob):
    """two Render_Pass calls == one call with twice the units (sample indices continue); Background_Color is added
    once per AA pixel-pass (integrators.adb:42)"""
    w, h = 64, 40
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.Background_Color = (0.125, 0.25, 0.5); rt.seed = 99
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    rt.Render_Pass()
    hdr = rt.Render_Pass(read_hdr=True)
    assert rt.GetSPP() == 16
    p = ob.make_params(w, h, ob.PT_MIS, passes=4, seed=99, background=(0.125, 0.25, 0.5))
    acc, _ = oscene.render(p)
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.all(hdr[..., 3] == 16.0)


def test_debug_render_palette(gscene, oscene, pkg, ob):
    """RT_DEBUG (ray_tracer.adb:208-238, 245-260): palette(matId rem 8), gamma, exact u32 frame"""
    w, h = 256, 192
    rt = pkg.Ray_Tracer()
    rt.Init_Render(pkg.RT_DEBUG); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    assert rt.Finished()
    acc, _ = oscene.render(ob.make_params(w, h, ob.RT_DEBUG))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, 1))


def test_one_shot_render_call(gscene, oscene, pkg, ob):
    """gcore_render (resize + units + read back into host buffers) == stateful calls"""
    w, h = 72, 40
    rt = pkg.Ray_Tracer(); rt.g_rend_type = pkg.PT_SHADOW; rt.seed = 5
    hdr = np.zeros((h, w, 4), np.float32); ldr = np.zeros((h, w), np.uint32)
    st = rt.render(w, h, 2, hdr, ldr)
    acc, rays = oscene.render(ob.make_params(w, h, ob.PT_SHADOW, passes=2, seed=5))
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.array_equal(ldr, ob.resolve(acc, 8))
    assert st.rays == rays and st.spp == 8 and st.kernel_launches == 3 and st.samples == w * h * 8


def test_full_size_properties(gscene, pkg):
    """BASELINE full sizes through size-independent properties (the oracle is too slow there): determinism, additivity
    of passes, energy sanity and LDR/HDR consistency at 2048x2048."""
    w = h = 2048
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 1; rt.seed = 0x5EED0004
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    a = rt.Render_Pass(read_hdr=True).copy()
    ldr_a = rt.screen_buffer.copy()
    rays_a = int(rt.last_stats.rays)
    rt.Resize_Viewport(w, h)
    b = rt.Render_Pass(read_hdr=True)
    assert np.array_equal(_bits(a), _bits(b)) and int(rt.last_stats.rays) == rays_a      # idempotent / deterministic
    assert np.isfinite(a).all() and (a[..., :3] >= 0).all()
    mean = a[..., :3].mean() / 4.0
    assert 0.02 < mean < 0.5
    # LDR is a pure function of HDR (ray_tracer.adb:281-291)
    v = np.sqrt(a[..., :3] * np.float32(0.25), dtype=np.float32)
    ref = np.floor(np.minimum(v, np.float32(1.0)) * np.float32(255.0) + np.float32(0.5)).astype(np.uint32)
    packed = ref[..., 0] | (ref[..., 1] << 8) | (ref[..., 2] << 16)
    assert (packed != ldr_a).mean() < 1e-5
    # rays per sample: every path traces at least the primary ray and at most 2 per bounce
    spp = 4
    assert w * h * spp <= rays_a <= w * h * spp * 16


def test_partition_independence(gscene, pkg):
    """tiles round-robin over `world` ranks; the gathered result is bit-identical to world == 1"""
    import torch
    w, h = 200, 136                                                 # not a multiple of 32: padded tiles
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 77
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    ref = rt.Render_Pass(read_hdr=True)
    ref_ldr = rt.screen_buffer.copy()
    for world in (2, 3, 8):
        bufs = []
        for rank in range(world):
            pkg.set_partition(rank, world)
            rt.Resize_Viewport(w, h)
            cfg = rt._cfg()
            assert pkg.lib().gcore_init_render(C.byref(cfg)) == 0
            st = pkg.RenderStats()
            assert pkg.lib().gcore_render_units(2, C.byref(st)) == 0
            ptr, nslots = pkg.accum_device_ptr()
            t = torch.empty(nslots * 4, dtype=torch.float32, device="cuda")
            pkg.synchronize()
            # device-to-device copy of this rank's slot-major accumulator (what the NCCL gather moves)
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr), C.c_size_t(nslots * 16), 3)
            bufs.append(t)
        allb = torch.cat(bufs)
        hdr = torch.empty(h * w * 4, dtype=torch.float32, device="cuda")
        ldr = torch.empty(h * w, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        assert pkg.lib().gcore_resolve_gathered_device(C.c_void_p(allb.data_ptr()), 8, C.c_void_p(hdr.data_ptr()),
                                                       C.c_void_p(ldr.data_ptr())) == 0
        pkg.synchronize()
        got = hdr.cpu().numpy().reshape(h, w, 4)
        assert np.array_equal(_bits(got[..., :3]), _bits(ref[..., :3])), "world=%d" % world
        assert np.array_equal(ldr.cpu().numpy().view(np.uint32).reshape(h, w), ref_ldr)
    pkg.set_partition(0, 1)


def test_checkpoint_resume(gscene, pkg, tmp_path):
    """2 passes + save + fresh viewport + load + 2 passes == 4 passes straight (sample indices continue at spp)"""
    w, h = 100, 60
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 4242
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    rt.Render_Pass(); ref = rt.Render_Pass(read_hdr=True).copy()
    assert rt.GetSPP() == 16
    rt.Resize_Viewport(w, h)
    rt.Render_Pass()
    ck = str(tmp_path / "ck.npz")
    rt.save_checkpoint(ck)
    rt2 = pkg.Ray_Tracer(); rt2.Threads_Num = 2; rt2.seed = 4242
    rt2.Init_Render(pkg.PT_MIS)
    rt2.load_checkpoint(ck)
    assert rt2.GetSPP() == 8
    got = rt2.Render_Pass(read_hdr=True)
    assert rt2.GetSPP() == 16 and np.array_equal(_bits(got), _bits(ref))


def test_c_harness_replays_test_adb(gscene, oscene, pkg, ob, tmp_path):
    """tests/c_harness/replay_test_adb.c: the call sequence of the reference's main program (test.adb:20-79) from plain C
    through dlopen -- no Python between the caller and the C ABI.  The BMP it writes equals the oracle's frame."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "replay_test_adb")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(root, "tests", "c_harness", "replay_test_adb.c"), "-ldl"], check=True)
    w, h, units, passes = 160, 120, 3, 2
    bmp = str(tmp_path / "ART_render.bmp")
    r = subprocess.run([exe, pkg.LIB_PATH, os.path.join(root, "data", "pyramid2.vsgf"), bmp, str(w), str(h), str(units), str(passes)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    spp, rays, _ms = r.stdout.split()
    acc, orays = oscene.render(ob.make_params(w, h, ob.PT_MIS, passes=units * passes, seed=0x5EED0003))
    assert int(spp) == 4 * units * passes and int(rays) == orays
    ref = str(tmp_path / "ref.bmp")
    ldr = np.ascontiguousarray(ob.resolve(acc, int(spp)).reshape(-1))
    assert ob.lib().ao_save_bmp(ref.encode(), ldr.ctypes.data, w, h) == 0
    assert open(bmp, "rb").read() == open(ref, "rb").read()


def test_argument_and_state_errors(pkg):
    """error behaviour of the ABI: negative codes + text, nothing throws (embree_connect.cpp exits on errors, :28-49)"""
    L = pkg.lib()
    st = pkg.RenderStats()
    cfg = pkg.make_config(max_depth=17)
    assert L.gcore_init_render(C.byref(cfg)) == -2 and "max_depth" in pkg.last_error()[1]
    cfg = pkg.make_config(render_type=9)
    assert L.gcore_init_render(C.byref(cfg)) == -2
    assert L.gcore_resize_viewport(0, 10) == -2 and L.gcore_resize_viewport(-5, 10) == -2
    assert L.gcore_render_units(-1, C.byref(st)) == -2
    assert L.gcore_set_partition(3, 2) == -2
    assert L.gcore_closest_hit_batch(-1, None, None) == -2
    assert L.gcore_closest_hit_batch(0, None, None) == 0
    assert L.gcore_load_frame(None, 4) == -2
    gc = pkg.GeometryCore(); gc.init_and_clear()
    assert gc.add_mesh_3f(np.zeros((0, 3), np.float32), np.zeros(0, np.int32)) == -1      # the reference returns an ambiguous 0
    gc.commit_scene()                                                                      # empty scene: nothing to build, nothing hit
    ok, h = gc.closest_hit([0, 0, 0], [0, 0, 1])
    assert not ok and h["primIndex"] == -1
    gc.destroy()


def test_tiny_viewports_and_zero_units(gscene, oscene, pkg, ob):
    for w, h in ((1, 1), (2, 3), (31, 1), (1, 33)):
        rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 17
        rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
        hdr = rt.Render_Pass(read_hdr=True)
        acc, rays = oscene.render(ob.make_params(w, h, ob.PT_MIS, passes=2, seed=17))
        assert np.array_equal(_bits(hdr[..., :3]), _bits(acc)) and int(rt.last_stats.rays) == rays
    st = pkg.RenderStats()
    assert pkg.lib().gcore_render_units(0, C.byref(st)) == 0 and st.rays == 0 and st.samples == 0


def test_long_renders_go_in_chunks_of_passes(oscene, ob, tmp_path):
    """the (pass, pixel) colour buffer is capped (2 GiB; C4 at 2048x2048 and C5 need 2+ launches per Render_Pass): with the cap
    forced down to 64 KB a 64x48 frame of 7 units takes 4 launches of 2 + 2 + 2 + 1 passes -- same bits as the oracle.
    Runs in a fresh process because the cap is read from the environment."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import __graft_entry__ as ge; pkg = ge.load_package(); "
        "pkg.set_device(0); sc = pkg.Scene(); sc.Init(); rt = pkg.Ray_Tracer(); rt.Threads_Num = 7; rt.seed = 99; "
        "rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(64, 48); hdr = rt.Render_Pass(read_hdr=True); "
        "np.save(%r, hdr); print(rt.last_stats.kernel_launches, rt.last_stats.rays)" % (root, str(tmp_path / "hdr.npy")))
    env = dict(os.environ, GCORE_COLBUF_CAP_KB="128")                     # 128 KB / (4096 slots x 16 B) = 2 passes per launch
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    launches, rays = (int(x) for x in r.stdout.split())
    hdr = np.load(str(tmp_path / "hdr.npy"))
    acc, orays = oscene.render(ob.make_params(64, 48, ob.PT_MIS, passes=7, seed=99))
    assert launches > 2 and rays == orays                                 # more than one (wavefront, fold) pair
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))


@pytest.mark.parametrize("tail_items", ["0", "777", "100000000"])
def test_launch_tail_goes_out_path_by_path(oscene, ob, tmp_path, tail_items):
    """the last (pass, pixel) items of a launch are handed out as four single sub-pixel paths each, so that the launch drains in one
    path's time (render_wave.cuh wave_regen; fold_kernel adds the four radiances in the reference's order).  Default: 2 x the resident
    path slots; here none / a handful / all of the items take that route -- the same bits as the oracle every time.  Fresh process:
    the setting is read from the environment when the work buffers are made."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import __graft_entry__ as ge; pkg = ge.load_package(); "
        "pkg.set_device(0); sc = pkg.Scene(); sc.Init(); rt = pkg.Ray_Tracer(); rt.Threads_Num = 5; rt.seed = 4099; "
        "rt.Background_Color = (0.03125, 0.0625, 0.125); rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(96, 80); "
        "hdr = rt.Render_Pass(read_hdr=True); np.save(%r, hdr); print(rt.last_stats.rays)" % (root, str(tmp_path / "hdr.npy")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, GCORE_TAIL_ITEMS=tail_items), timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    hdr = np.load(str(tmp_path / "hdr.npy"))
    acc, orays = oscene.render(ob.make_params(96, 80, ob.PT_MIS, passes=5, seed=4099, background=(0.03125, 0.0625, 0.125)))
    assert int(r.stdout.split()[-1]) == orays
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))


@pytest.mark.parametrize("name,rtype,depth,units,w,h", [
    ("C1", "PT_MIS", 8, 2, 1024, 768),        # the reference's default viewport and integrator, 8 of its 112 spp per pass
    ("C3", "PT_SHADOW", 1, 4, 1024, 1024),    # BASELINE config C3 in full: direct lighting, 16 spp
    ("C4", "PT_MIS", 8, 1, 2048, 2048),       # BASELINE config C4 at full resolution, 4 of its 256 spp
])
def test_baseline_configs_at_full_resolution_bit_exact(gscene, oscene, pkg, ob, name, rtype, depth, units, w, h):
    """the BASELINE viewports in full (the multi-threaded oracle renders them in a second or two): every accumulator bit, the ray
    count and the LDR frame"""
    rt = pkg.Ray_Tracer()
    rt.Threads_Num = units; rt.Max_Trace_Depth = depth; rt.seed = 0x5EED0003 if name == "C3" else 0x5EED0004
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    hdr = rt.Render_Pass(read_hdr=True)
    acc, rays = oscene.render(ob.make_params(w, h, getattr(ob, rtype), max_depth=depth, passes=units, seed=rt.seed))
    assert int(rt.last_stats.rays) == rays
    assert np.array_equal(_bits(hdr[..., :3]), _bits(acc))
    assert np.array_equal(rt.screen_buffer, ob.resolve(acc, 4 * units))
