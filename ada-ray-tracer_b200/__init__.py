"""ada-ray-tracer_b200 -- host-side binding of libgcore_b200.so (include/gcore.h).

The classes mirror the reference's Ada packages for the hot path, with the same names and argument meaning:

  Scene       scene.ads:17-38            Init, Destroy, Find_Closest_Hit, Material_At, Light_At, Camera_At
  Ray_Tracer  ray_tracer.ads:20-48       width/height/Threads_Num/Anti_Aliasing_On/Max_Trace_Depth/Background_Color,
                                         Init_Render, Resize_Viewport, Render_Pass, GetSPP, Finished, screen_buffer
  Bitmap      bitmap.ads:10-25           Image, Init, SaveBMP

Everything that computes goes through the C ABI into CUDA kernels; there is no Python or CPU implementation of the
path here, and loading fails loudly if the shared library has not been built.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libgcore_b200.so")      # the product library (bit-exact numerics); experiment builds are loaded by profiles/tools/bench_variant.py
FAST_LIB_PATH = os.path.join(_HERE, "libgcore_b200_fast.so")   # the NUMERICS=fast experiment build of the same sources (csrc/Makefile)
NUMERICS = "exact"
DEFAULT_VSGF = os.path.join(ROOT, "data", "pyramid2.vsgf")

# ---- POD layouts of include/gcore.h -------------------------------------------------------------------------------
sphere_dt = np.dtype([("pos", "<f4", 3), ("r", "<f4"), ("mat", "<i4")])
cornell_dt = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("mat_indices", "<i4", 6), ("normals", "<f4", 18)])
flatlight_dt = np.dtype([("boxMin", "<f4", 3), ("boxMax", "<f4", 3), ("normal", "<f4", 3), ("intensity", "<f4", 3),
                         ("surfaceArea", "<f4")])
material_dt = np.dtype([("type", "<i4"), ("color", "<f4", 3), ("color2", "<f4", 3), ("param", "<f4"), ("light", "<i4")])
light_dt = np.dtype([("shape", "<i4"), ("boxMin", "<f4", 3), ("boxMax", "<f4", 3), ("normal", "<f4", 3),
                     ("center", "<f4", 3), ("radius", "<f4"), ("intensity", "<f4", 3), ("surfaceArea", "<f4")])
camera_dt = np.dtype([("pos", "<f4", 3), ("lookAt", "<f4", 3), ("up", "<f4", 3), ("matrix", "<f4", 16)])
hit_dt = np.dtype([("is_hit", "<i4"), ("prim_type", "<i4"), ("prim_index", "<i4"), ("matId", "<i4"), ("mat", "<i4"),
                   ("t", "<f4"), ("normal", "<f4", 3), ("tx", "<f4"), ("ty", "<f4")])
ray_dt = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3)])
hitcpp_dt = np.dtype([("primIndex", "<i4"), ("geomIndex", "<i4"), ("instIndex", "<i4"), ("t", "<f4"),
                      ("normal", "<f4", 3), ("texCoord", "<f4", 2)])

RT_DEBUG, RT_WHITTED, PT_STUPID, PT_SHADOW, PT_MIS = range(5)          # ray_tracer.ads:40
MESH_REFERENCE_BF, MESH_NEAREST = 0, 1
MAT_NULL, MAT_LIGHT, MAT_LAMBERT, MAT_MIRROR, MAT_GLASS, MAT_PHONG = range(6)
LIGHT_RECT, LIGHT_SPHERE = 0, 1


class RenderConfig(C.Structure):
    _fields_ = [("render_type", C.c_int32), ("anti_aliasing", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint32), ("background", C.c_float * 3)]


class RenderStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("mesh_queries", C.c_uint64), ("bounces", C.c_uint64), ("reserved0", C.c_uint64),
                ("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("spp", C.c_int32), ("kernel_launches", C.c_int32)]


class GCoreError(RuntimeError):
    pass


EXPORTS = [
    "gcore_init_and_clear", "gcore_destroy", "gcore_add_mesh_3f", "gcore_instance_meshes", "gcore_commit_scene",
    "gcore_closest_hit", "gcore_set_analytic_scene", "gcore_set_materials", "gcore_set_light", "gcore_set_camera",
    "gcore_set_mesh", "gcore_set_mesh_vsgf", "gcore_scene_init_cornell", "gcore_get_mesh", "gcore_closest_hit_batch",
    "gcore_bvh_closest_hit_batch", "gcore_set_device", "gcore_set_partition", "gcore_init_render",
    "gcore_resize_viewport", "gcore_render_units", "gcore_get_counters", "gcore_get_spp", "gcore_read_frame", "gcore_render",
    "gcore_primary_hits", "gcore_accum_device_ptr", "gcore_resolve_device", "gcore_resolve_gathered_device",
    "gcore_stream", "gcore_synchronize", "gcore_last_error", "gcore_bvh_info", "gcore_kernel_regs", "gcore_bvh_bench", "gcore_load_frame", "gcore_set_stats_level", "gcore_get_bvh_counters",
    "gcore_set_gpu_count", "gcore_get_gpu_count", "gcore_device_kernel_ms", "gcore_set_seam_two_sided", "gcore_set_mesh_material",
    "gcore_set_mesh_normals_3f", "gcore_use_committed_scene", "gcore_get_committed_scene", "gcore_bvh_occluded_batch", "gcore_seam_stats", "gcore_seam_server_stats",
]

_lib = None


def select_numerics(mode):
    """"exact" (default, the product): libgcore_b200.so, bit-identical to the reference arithmetic.  "fast": libgcore_b200_fast.so, an
    EXPERIMENT build of the same sources (FMA contraction, approximate division / square root): 18 % faster, but it shifts image means by
    up to ~1.6 % and does not meet SURVEY 8d's block-RMSE acceptance (tests/test_gpu_fast_numerics.py says why) -- for measurements only.
    One library per process: call before anything else of this package."""
    global LIB_PATH, NUMERICS
    if _lib is not None:
        raise GCoreError("select_numerics must be called before the library is loaded")
    if mode not in ("exact", "fast"):
        raise GCoreError("numerics mode must be 'exact' or 'fast'")
    NUMERICS = mode
    if mode == "fast":
        LIB_PATH = FAST_LIB_PATH


def lib():
    """Load libgcore_b200.so (built in-tree by `make -C ada-ray-tracer_b200/csrc` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GCoreError("libgcore_b200.so is not built (%s); run __graft_entry__.build() -- there is no fallback path"
                         % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.gcore_init_and_clear.restype = None
    L.gcore_destroy.restype = None
    L.gcore_add_mesh_3f.argtypes = [vp, i32, vp, i32]
    L.gcore_instance_meshes.argtypes = [i32, vp, i32]
    L.gcore_instance_meshes.restype = None
    L.gcore_commit_scene.restype = None
    L.gcore_closest_hit.argtypes = [vp, vp, f32, f32, vp]
    L.gcore_closest_hit.restype = C.c_bool
    L.gcore_set_analytic_scene.argtypes = [vp, i32, vp, vp]
    L.gcore_set_materials.argtypes = [vp, i32]
    L.gcore_set_light.argtypes = [vp]
    L.gcore_set_camera.argtypes = [vp]
    L.gcore_set_mesh.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    L.gcore_set_mesh_vsgf.argtypes = [C.c_char_p, vp, i32]
    L.gcore_scene_init_cornell.argtypes = [C.c_char_p]
    L.gcore_get_mesh.argtypes = [vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
    L.gcore_closest_hit_batch.argtypes = [i32, vp, vp]
    L.gcore_bvh_closest_hit_batch.argtypes = [i32, vp, f32, f32, vp, vp]
    L.gcore_set_device.argtypes = [i32]
    L.gcore_set_partition.argtypes = [i32, i32]
    L.gcore_init_render.argtypes = [C.POINTER(RenderConfig)]
    L.gcore_resize_viewport.argtypes = [i32, i32]
    L.gcore_render_units.argtypes = [i32, C.POINTER(RenderStats)]
    L.gcore_get_counters.argtypes = [vp]
    L.gcore_get_bvh_counters.argtypes = [vp]
    L.gcore_read_frame.argtypes = [vp, vp]
    L.gcore_load_frame.argtypes = [vp, i32]
    L.gcore_set_stats_level.argtypes = [i32]
    L.gcore_render.argtypes = [C.POINTER(RenderConfig), i32, i32, i32, vp, vp, C.POINTER(RenderStats)]
    L.gcore_primary_hits.argtypes = [i32, i32, vp]
    L.gcore_accum_device_ptr.argtypes = [C.POINTER(vp), C.POINTER(C.c_int64)]
    L.gcore_resolve_device.argtypes = [vp, vp]
    L.gcore_resolve_gathered_device.argtypes = [vp, i32, vp, vp]
    L.gcore_stream.argtypes = [C.POINTER(vp)]
    L.gcore_last_error.argtypes = [C.c_char_p, i32]
    L.gcore_bvh_info.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(f32)]
    L.gcore_kernel_regs.argtypes = [i32]
    L.gcore_bvh_bench.argtypes = [i32, vp, f32, f32, i32, i32, C.POINTER(f32), vp]
    L.gcore_set_gpu_count.argtypes = [i32]
    L.gcore_device_kernel_ms.argtypes = [vp, i32]
    L.gcore_set_seam_two_sided.argtypes = [i32]
    L.gcore_set_mesh_material.argtypes = [i32, i32]
    L.gcore_set_mesh_normals_3f.argtypes = [i32, vp, i32]
    L.gcore_use_committed_scene.argtypes = [i32]
    L.gcore_get_committed_scene.argtypes = [vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
    L.gcore_bvh_occluded_batch.argtypes = [i32, vp, f32, f32, vp]
    L.gcore_seam_stats.argtypes = [vp]
    L.gcore_seam_server_stats.argtypes = [vp]
    _lib = L
    return L


def last_error():
    buf = C.create_string_buffer(512)
    code = lib().gcore_last_error(buf, 512)
    return code, buf.value.decode(errors="replace")


def _ck(rc):
    if rc != 0:
        code, text = last_error()
        raise GCoreError("gcore error %d: %s" % (rc, text))


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def make_config(render_type=PT_MIS, anti_aliasing=True, max_depth=8, seed=0x5EED0003, background=(0.0, 0.0, 0.0)):
    cfg = RenderConfig(int(render_type), 1 if anti_aliasing else 0, int(max_depth), int(seed))
    cfg.background = (C.c_float * 3)(*background)
    return cfg


# ======================================================================================================================
# Scene  (scene.ads:17-38)
# ======================================================================================================================
class Scene:
    """Render_Scene of the default ("internal_ada") build, held on the device by the render core."""

    def __init__(self):
        self.L = lib()

    def Init(self, a_path=None, vsgf=DEFAULT_VSGF):
        """Scene.Init (scene.adb:24-27 -> Init_Cornell_Box :89-217).  a_path is ignored like in the reference."""
        _ck(self.L.gcore_scene_init_cornell(vsgf.encode()))

    def Destroy(self):
        """Scene.Destroy (scene.adb:29-32): nothing to do in the analytic build."""

    # --- explicit description (what Init_Cornell_Box hard-codes) ---------------------------------------------------
    def set_analytic(self, spheres, cornell, flatlight=None):
        spheres = np.ascontiguousarray(spheres, sphere_dt)
        cornell = np.ascontiguousarray(cornell, cornell_dt)
        fl = None if flatlight is None else np.ascontiguousarray(flatlight, flatlight_dt)
        _ck(self.L.gcore_set_analytic_scene(_p(spheres), len(spheres), _p(cornell), _p(fl)))

    def set_materials(self, mats):
        mats = np.ascontiguousarray(mats, material_dt)
        _ck(self.L.gcore_set_materials(_p(mats), len(mats)))

    def set_light(self, light):
        light = np.ascontiguousarray(light, light_dt)
        _ck(self.L.gcore_set_light(_p(light)))

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam, camera_dt)
        _ck(self.L.gcore_set_camera(_p(cam)))

    def set_mesh(self, pos, nrm, tris, bbox=None, mesh_mode=MESH_REFERENCE_BF):
        pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        bb = None if bbox is None else np.ascontiguousarray(bbox, np.float32)
        _ck(self.L.gcore_set_mesh(_p(pos), _p(nrm), len(pos), _p(tris), len(tris), _p(bb), mesh_mode))

    def load_vsgf(self, path, matrix16, mesh_mode=MESH_REFERENCE_BF):
        """Geometry.LoadMeshFromVSGF (geometry.adb:499-609)."""
        m = np.ascontiguousarray(matrix16, np.float32).reshape(16)
        _ck(self.L.gcore_set_mesh_vsgf(path.encode(), _p(m), mesh_mode))

    def get_mesh(self):
        nv, nt = C.c_int(), C.c_int()
        _ck(self.L.gcore_get_mesh(None, None, None, None, C.byref(nv), C.byref(nt)))
        pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nv.value, 3), np.float32)
        tris = np.zeros((nt.value, 3), np.int32); bbox = np.zeros(6, np.float32)
        _ck(self.L.gcore_get_mesh(_p(pos), _p(nrm), _p(tris), _p(bbox), C.byref(nv), C.byref(nt)))
        return pos, nrm, tris, bbox

    # --- queries ---------------------------------------------------------------------------------------------------
    def Find_Closest_Hit(self, rays):
        """Scene.Find_Closest_Hit (scene.adb:56-86) for an array of rays (ray_dt) -> array of hit_dt."""
        rays = np.ascontiguousarray(rays, ray_dt)
        hits = np.zeros(len(rays), hit_dt)
        _ck(self.L.gcore_closest_hit_batch(len(rays), _p(rays), _p(hits)))
        return hits

    def primary_hits(self, width, height):
        hits = np.zeros(width * height, hit_dt)
        _ck(self.L.gcore_primary_hits(width, height, _p(hits)))
        return hits.reshape(height, width)


# ======================================================================================================================
# legacy geometry core (cpp/embree_connect.cpp as imported by scene_hydra_embree.adb:37-66)
# ======================================================================================================================
class GeometryCore:
    def __init__(self):
        self.L = lib()

    def init_and_clear(self):
        self.L.gcore_init_and_clear()

    def destroy(self):
        self.L.gcore_destroy()

    def add_mesh_3f(self, verts, indices):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, np.int32).reshape(-1)
        return self.L.gcore_add_mesh_3f(_p(verts), len(verts), _p(indices), len(indices))

    def instance_meshes(self, mesh_id, matrices16):
        m = np.ascontiguousarray(matrices16, np.float32).reshape(-1, 16)
        self.L.gcore_instance_meshes(mesh_id, _p(m), len(m))

    def commit_scene(self):
        self.L.gcore_commit_scene()
        code, text = last_error()
        if code != 0:
            raise GCoreError("gcore_commit_scene: %s" % text)

    def closest_hit(self, pos, direction, t_near=0.0, t_far=100000.0):
        p = np.ascontiguousarray(pos, np.float32); d = np.ascontiguousarray(direction, np.float32)
        h = np.zeros(1, hitcpp_dt); h["primIndex"] = -1; h["geomIndex"] = -1; h["instIndex"] = -1
        found = self.L.gcore_closest_hit(_p(p), _p(d), t_near, t_far, _p(h))
        return bool(found), h[0]

    def closest_hit_batch(self, rays, t_near=0.0, t_far=100000.0):
        rays = np.ascontiguousarray(rays, ray_dt)
        hits = np.zeros(len(rays), hitcpp_dt); hits["primIndex"] = -1
        found = np.zeros(len(rays), np.int32)
        _ck(self.L.gcore_bvh_closest_hit_batch(len(rays), _p(rays), t_near, t_far, _p(hits), _p(found)))
        return found, hits

    def set_mesh_material(self, mesh_id, mat_id):
        _ck(self.L.gcore_set_mesh_material(mesh_id, mat_id))

    def set_mesh_normals_3f(self, mesh_id, normals):
        n = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        _ck(self.L.gcore_set_mesh_normals_3f(mesh_id, _p(n), len(n)))

    def use_committed_scene(self, enable=True):
        """the integrators trace the committed scene (meshes + instances + per-mesh materials) instead of the Scene.set_mesh mesh"""
        _ck(self.L.gcore_use_committed_scene(1 if enable else 0))

    def set_two_sided(self, enable):
        _ck(self.L.gcore_set_seam_two_sided(1 if enable else 0))

    def committed_scene(self):
        """-> (pos[nv,3], tris[nt,3], normals[nt,3,3], mats[nt]) of the flattened world-space scene the devices hold"""
        nv, nt = C.c_int(), C.c_int()
        _ck(self.L.gcore_get_committed_scene(None, None, None, None, C.byref(nv), C.byref(nt)))
        pos = np.zeros((nv.value, 3), np.float32); tris = np.zeros((nt.value, 3), np.int32)
        n9 = np.zeros((nt.value, 3, 3), np.float32); mats = np.zeros(nt.value, np.int32)
        _ck(self.L.gcore_get_committed_scene(_p(pos), _p(tris), _p(n9), _p(mats), C.byref(nv), C.byref(nt)))
        return pos, tris, n9, mats

    def occluded_batch(self, rays, t_near=0.0, t_far=100000.0):
        rays = np.ascontiguousarray(rays, ray_dt)
        occ = np.zeros(len(rays), np.int32)
        _ck(self.L.gcore_bvh_occluded_batch(len(rays), _p(rays), t_near, t_far, _p(occ)))
        return occ

    def bvh_info(self):
        a, b, c, ms = C.c_int64(), C.c_int64(), C.c_int64(), C.c_float()
        self.L.gcore_bvh_info(C.byref(a), C.byref(b), C.byref(c), C.byref(ms))
        return {"ntris": a.value, "nnodes": b.value, "bytes": c.value, "build_ms": ms.value}


# ======================================================================================================================
# Ray_Tracer  (ray_tracer.ads:20-48)
# ======================================================================================================================
class Ray_Tracer:
    """Package Ray_Tracer: configuration variables, Init_Render, Resize_Viewport, Render_Pass, GetSPP, Finished."""

    def __init__(self):
        self.L = lib()
        self.width = 1024                       # ray_tracer.ads:20
        self.height = 768                       # :21
        self.Threads_Num = 14 * 2               # :23 -- DoPass units per Render_Pass
        self.Anti_Aliasing_On = True            # :24
        self.Max_Trace_Depth = 8                # :25
        self.Background_Color = (0.0, 0.0, 0.0)  # :27
        self.seed = 0x5EED0003
        self.g_rend_type = PT_MIS               # :108
        self.g_finish = False                   # :109
        self.screen_buffer = None               # :37  (exposed row-major [y, x], see Render_Pass)
        self.last_stats = RenderStats()

    def _cfg(self):
        return make_config(self.g_rend_type, self.Anti_Aliasing_On, self.Max_Trace_Depth, self.seed, self.Background_Color)

    def Init_Render(self, a_rendType):
        """ray_tracer.adb:197-200"""
        self.g_rend_type = int(a_rendType)
        cfg = self._cfg()
        _ck(self.L.gcore_init_render(C.byref(cfg)))

    def Resize_Viewport(self, size_x, size_y):
        """ray_tracer.adb:297-320"""
        self.width, self.height = int(size_x), int(size_y)
        _ck(self.L.gcore_resize_viewport(self.width, self.height))
        self.screen_buffer = np.zeros((self.height, self.width), np.uint32)

    def Render_Pass(self, read_hdr=False):
        """ray_tracer.adb:240-293: Threads_Num DoPass units, then resolve into screen_buffer.

        screen_buffer is returned row-major [y, x] (the reference's screen_buffer(x, y) transposed the way test.adb:63-67
        copies it into Bitmap.Image.data)."""
        cfg = self._cfg()
        _ck(self.L.gcore_init_render(C.byref(cfg)))
        units = 1 if self.g_rend_type in (RT_DEBUG, RT_WHITTED) else self.Threads_Num
        _ck(self.L.gcore_render_units(units, C.byref(self.last_stats)))
        hdr = np.zeros((self.height, self.width, 4), np.float32) if read_hdr else None
        _ck(self.L.gcore_read_frame(_p(hdr), _p(self.screen_buffer)))
        if self.g_rend_type in (RT_DEBUG, RT_WHITTED):
            self.g_finish = True                # ray_tracer.adb:259
        return hdr

    def GetSPP(self):
        return self.L.gcore_get_spp()

    def save_checkpoint(self, path):
        """progressive-render checkpoint: HDR accumulator + spp (np.savez)"""
        hdr = np.zeros((self.height, self.width, 4), np.float32)
        _ck(self.L.gcore_read_frame(_p(hdr), None))
        np.savez(path, hdr=hdr, spp=self.GetSPP(), width=self.width, height=self.height)

    def load_checkpoint(self, path):
        z = np.load(path)
        self.Resize_Viewport(int(z["width"]), int(z["height"]))
        hdr = np.ascontiguousarray(z["hdr"], np.float32)
        _ck(self.L.gcore_load_frame(_p(hdr), int(z["spp"])))

    def Finished(self):
        return self.g_finish

    # --- one-shot end-to-end call --------------------------------------------------------------------------------
    def render(self, width, height, units, hdr=None, ldr=None):
        """gcore_render: resize + `units` DoPass units + read back into the given host buffers."""
        cfg = self._cfg()
        st = RenderStats()
        _ck(self.L.gcore_render(C.byref(cfg), width, height, units, _p(hdr), _p(ldr), C.byref(st)))
        self.width, self.height = width, height
        self.last_stats = st
        return st


# ======================================================================================================================
# Bitmap  (bitmap.ads:10-25, bitmap.adb:31-85)
# ======================================================================================================================
class Bitmap:
    class Image:
        def __init__(self):
            self.width = 0
            self.height = 0
            self.data = None                    # PixelData: Unsigned_32, data(y*width + x)

    @staticmethod
    def Init(im, w, h):
        im.width, im.height = int(w), int(h)
        im.data = np.zeros(im.width * im.height, np.uint32)

    @staticmethod
    def Delete(im):
        im.width = im.height = 0
        im.data = None

    @staticmethod
    def SaveBMP(im, a_fileName):
        """24-bit BMP, bottom-up, B,G,R bytes, no row padding -- byte for byte what bitmap.adb:31-85 writes."""
        w, h = im.width, im.height
        hdr = np.zeros(54, np.uint8)
        hdr[0:2] = (0x42, 0x4D)
        hdr[2:6] = np.frombuffer(np.uint32(14 + 40 + w * h * 3).tobytes(), np.uint8)
        hdr[10:14] = np.frombuffer(np.uint32(54).tobytes(), np.uint8)
        hdr[14:18] = np.frombuffer(np.uint32(40).tobytes(), np.uint8)
        hdr[18:22] = np.frombuffer(np.uint32(w).tobytes(), np.uint8)
        hdr[22:26] = np.frombuffer(np.uint32(h).tobytes(), np.uint8)
        hdr[26:28] = (1, 0)
        hdr[28:30] = (24, 0)
        px = np.asarray(im.data, np.uint32)
        body = np.empty((px.size, 3), np.uint8)
        body[:, 0] = (px >> 16) & 255           # Pixel.r field  <- bits 16..23 (blue of the packed colour)
        body[:, 1] = (px >> 8) & 255
        body[:, 2] = px & 255                   # Pixel.b field  <- bits 0..7   (red of the packed colour)
        with open(a_fileName, "wb") as f:
            f.write(hdr.tobytes())
            f.write(body.tobytes())


def bvh_bench(rays, which=1, repeat=5, t_near=0.0, t_far=1000000.0):
    """traversal micro-benchmark: -> dict(ms, mrays_s, nodes_per_ray, tris_per_ray, bytes_per_ray, hit_rate)"""
    rays = np.ascontiguousarray(rays, ray_dt)
    ms = C.c_float(); st = np.zeros(3, np.uint64)
    _ck(lib().gcore_bvh_bench(len(rays), _p(rays), t_near, t_far, which, repeat, C.byref(ms), _p(st)))
    n = float(len(rays))
    nodes, tris = st[0] / n, st[1] / n
    return {"ms": ms.value, "mrays_s": n / (ms.value * 1e-3) / 1e6, "nodes_per_ray": float(nodes), "tris_per_ray": float(tris),
            "bytes_per_ray": float(64.0 * nodes + 48.0 * tris + 24 + 40), "hit_rate": float(st[2] / n)}


def set_stats_level(level):
    """1: fast counters; 2: mesh_queries = rays entering the reference's mesh bbox (for the roofline's flop count)"""
    _ck(lib().gcore_set_stats_level(int(level)))


def set_device(device):
    _ck(lib().gcore_set_device(int(device)))


def set_partition(rank, world):
    _ck(lib().gcore_set_partition(int(rank), int(world)))


def set_gpu_count(n):
    """multi-GPU inside this one process: devices [device, device + n) split the tiles, gcore_read_frame gathers with NCCL"""
    _ck(lib().gcore_set_gpu_count(int(n)))


def get_gpu_count():
    return lib().gcore_get_gpu_count()


def device_kernel_ms(n=None):
    n = n or get_gpu_count()
    out = np.zeros(n, np.float32)
    _ck(lib().gcore_device_kernel_ms(_p(out), n))
    return out


def accum_device_ptr():
    p, n = C.c_void_p(), C.c_int64()
    _ck(lib().gcore_accum_device_ptr(C.byref(p), C.byref(n)))
    return p.value, n.value


def stream_handle():
    p = C.c_void_p()
    _ck(lib().gcore_stream(C.byref(p)))
    return p.value


def get_counters():
    """(rays traced, mesh_queries, bounces, LBVH nodes visited) accumulated since the last Resize_Viewport"""
    c = np.zeros(4, np.uint64)
    _ck(lib().gcore_get_counters(_p(c)))
    return int(c[0]), int(c[1]), int(c[2]), int(c[3])


def get_bvh_counters():
    """(LBVH nodes visited, triangle slots tested, traversals started) by the render kernel since the last Resize_Viewport"""
    c = np.zeros(3, np.uint64)
    _ck(lib().gcore_get_bvh_counters(_p(c)))
    return int(c[0]), int(c[1]), int(c[2])


def synchronize():
    _ck(lib().gcore_synchronize())


# ======================================================================================================================
# Render_Pass over the GPUs of one box: tiles split round-robin, accumulators gathered with NCCL
# ======================================================================================================================
class _DevicePtr:
    """zero-copy view of library-owned device memory for torch (CUDA array interface)"""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class Tiled_Ray_Tracer:
    """One instance per process (= per GPU).  Mirrors Ray_Tracer's call sequence (Init_Render / Resize_Viewport /
    Render_Pass); Render_Pass leaves the resolved frame on rank 0.  torch supplies streams and torch.distributed only."""

    def __init__(self, rt, local_rank=0, group=None):
        import torch
        import torch.distributed as dist
        from . import tiles
        self.torch, self.tiles = torch, tiles
        self.rt = rt
        self.L = lib()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", local_rank)
        set_device(local_rank)
        set_partition(self.rank, self.world)
        # torch sees the library's stream: NCCL and copies are ordered with the kernels without host syncs
        self.stream = torch.cuda.ExternalStream(stream_handle(), device=self.device)
        self.gather = None
        self.d_ldr = self.d_hdr = self.local = None

    def Resize_Viewport(self, size_x, size_y):
        torch = self.torch
        self.rt.Resize_Viewport(size_x, size_y)
        ptr, nslots = accum_device_ptr()
        assert nslots == self.tiles.local_slots(size_x, size_y, self.world)
        with torch.cuda.stream(self.stream):
            self.local = torch.as_tensor(_DevicePtr(ptr, nslots * 4), device=self.device)
            if self.gather is None or self.gather.nslots != nslots:
                self.gather = self.tiles.FrameGather(nslots, 4, self.device, torch.float32, self.group) if self.world > 1 else None
            if self.rank == 0 and (self.d_ldr is None or self.d_ldr.numel() != size_x * size_y):
                self.d_ldr = torch.empty(size_x * size_y, dtype=torch.int32, device=self.device)
                self.d_hdr = None

    def Render_Pass(self, host_ldr=None, want_hdr=False, events=None):
        """Threads_Num DoPass units on every rank's tiles, gather, resolve on rank 0 (into self.d_ldr / self.d_hdr, device
        memory); host_ldr: optional pinned int32 tensor that receives the frame on rank 0 (asynchronous copy on the
        library's stream; synchronize before reading); events: optional 3 torch.cuda.Event recorded on the library's stream."""
        torch, rt = self.torch, self.rt
        cfg = rt._cfg()
        _ck(self.L.gcore_init_render(C.byref(cfg)))
        if events:
            events[0].record(self.stream)
        _ck(self.L.gcore_render_units(1 if rt.g_rend_type in (RT_DEBUG, RT_WHITTED) else rt.Threads_Num, None))
        if events:
            events[1].record(self.stream)                # events[0..1]: the render kernels alone
        npx = rt.width * rt.height
        with torch.cuda.stream(self.stream):
            if self.rank == 0 and want_hdr and self.d_hdr is None:
                self.d_hdr = torch.empty(npx * 4, dtype=torch.float32, device=self.device)
            hdr_p = C.c_void_p(self.d_hdr.data_ptr()) if (self.rank == 0 and want_hdr) else None
            if self.world == 1:
                _ck(self.L.gcore_resolve_device(hdr_p, C.c_void_p(self.d_ldr.data_ptr())))
            else:
                allb = self.gather.gather(self.local)
                if self.rank == 0:
                    spp = 1 if rt.g_rend_type in (RT_DEBUG, RT_WHITTED) else rt.GetSPP()      # the palette is resolved without 1/spp
                    _ck(self.L.gcore_resolve_gathered_device(C.c_void_p(allb.data_ptr()), spp, hdr_p, C.c_void_p(self.d_ldr.data_ptr())))
            if self.rank == 0 and host_ldr is not None:
                host_ldr.copy_(self.d_ldr, non_blocking=True)
        if events:
            events[2].record(self.stream)                # events[0..2]: the whole pass on this rank
        return 2 + (1 if self.rank == 0 else 0)          # kernels launched by this rank: wavefront + fold (+ resolve)
