"""ctypes binding of the CPU oracle (oracle/art_oracle.h).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libart_oracle.so")
VSGF = os.path.join(ROOT, "data", "pyramid2.vsgf")

# --- POD layouts shared (by byte layout) between oracle/art_oracle.h and include/gcore.h ------------
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

RT_DEBUG, RT_WHITTED, PT_STUPID, PT_SHADOW, PT_MIS = range(5)
MESH_REFERENCE_BF, MESH_NEAREST = 0, 1
MAT_NULL, MAT_LIGHT, MAT_LAMBERT, MAT_MIRROR, MAT_GLASS, MAT_PHONG = range(6)


class Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("render_type", C.c_int32), ("anti_aliasing", C.c_int32),
                ("max_depth", C.c_int32), ("passes", C.c_int32), ("seed", C.c_uint32), ("first_sample", C.c_int32),
                ("background", C.c_float * 3), ("mesh_mode", C.c_int32), ("threads", C.c_int32),
                ("y0", C.c_int32), ("y1", C.c_int32)]


def make_params(width, height, render_type=PT_MIS, aa=1, max_depth=8, passes=1, seed=0x5EED0003, first_sample=0,
                background=(0.0, 0.0, 0.0), mesh_mode=MESH_REFERENCE_BF, threads=0, y0=0, y1=0):
    p = Params(width, height, render_type, aa, max_depth, passes, seed, first_sample)
    p.background = (C.c_float * 3)(*background)
    p.mesh_mode, p.threads, p.y0, p.y1 = mesh_mode, threads, y0, y1
    return p


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        L.ao_scene_new.restype = vp
        L.ao_scene_free.argtypes = [vp]
        L.ao_scene_init_cornell.argtypes = [vp, C.c_char_p]
        L.ao_set_spheres.argtypes = [vp, vp, C.c_int]
        L.ao_set_cornell.argtypes = [vp, vp]
        L.ao_set_flatlight.argtypes = [vp, vp]
        L.ao_set_materials.argtypes = [vp, vp, C.c_int]
        L.ao_set_light.argtypes = [vp, vp]
        L.ao_set_camera.argtypes = [vp, vp]
        L.ao_set_mesh.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int]
        L.ao_load_vsgf.argtypes = [vp, C.c_char_p, vp]
        L.ao_set_mesh_tri_materials.argtypes = [vp, vp, C.c_int]
        L.ao_bruteforce_closest_two_sided.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp]
        L.ao_get_counts.argtypes = [vp] + [C.POINTER(C.c_int)] * 4
        for n in ("spheres", "cornell", "materials", "light", "camera"):
            getattr(L, "ao_get_" + n).argtypes = [vp, vp]
        L.ao_get_mesh.argtypes = [vp, vp, vp, vp, vp]
        L.ao_closest_hit_batch.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.ao_eye_ray.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.ao_primary_hits.argtypes = [vp, C.POINTER(Params), vp]
        L.ao_compute_shadow.argtypes = [vp, C.c_int, vp, vp]
        L.ao_render.argtypes = [vp, C.POINTER(Params), vp]
        L.ao_render.restype = C.c_uint64
        L.ao_render_task_lock.argtypes = [vp, C.POINTER(Params), C.c_int, vp]
        L.ao_render_task_lock.restype = C.c_uint64
        L.ao_resolve.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ao_save_bmp.argtypes = [C.c_char_p, vp, C.c_int, C.c_int]
        L.ao_philox4x32_10.argtypes = [vp, vp, vp]
        L.ao_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ao_pow.argtypes = [C.c_float, C.c_float]
        L.ao_pow.restype = C.c_float
        L.ao_use_libm.argtypes = [C.c_int]
        L.ao_bvh_build.argtypes = [vp, C.c_int, vp, C.c_int]
        L.ao_bvh_build.restype = vp
        L.ao_bvh_free.argtypes = [vp]
        L.ao_bvh_closest.argtypes = [vp, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp, C.c_int]
        L.ao_bruteforce_closest.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp]
        L.ao_sincos_batch.argtypes = [C.c_int, vp, vp, vp]
        L.ao_expon_batch.argtypes = [C.c_int, vp, vp, vp]
        L.ao_light_sample_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.ao_light_eval_pdf_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.ao_mat_sample_batch.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.ao_mat_eval_batch.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleScene:
    """Scene.Render_Scene of the oracle."""

    def __init__(self, cornell=True, vsgf=VSGF):
        self.L = lib()
        self.h = self.L.ao_scene_new()
        if cornell:
            rc = self.L.ao_scene_init_cornell(self.h, vsgf.encode())
            if rc != 0:
                raise RuntimeError("ao_scene_init_cornell failed: %d" % rc)

    def close(self):
        if self.h:
            self.L.ao_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- description -----------------------------------------------------------------------------
    def counts(self):
        v = [C.c_int() for _ in range(4)]
        self.L.ao_get_counts(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def describe(self):
        ns, nm, nv, nt = self.counts()
        d = {}
        d["spheres"] = np.zeros(ns, sphere_dt); self.L.ao_get_spheres(self.h, _p(d["spheres"]))
        d["cornell"] = np.zeros(1, cornell_dt); self.L.ao_get_cornell(self.h, _p(d["cornell"]))
        d["materials"] = np.zeros(nm, material_dt); self.L.ao_get_materials(self.h, _p(d["materials"]))
        d["light"] = np.zeros(1, light_dt); self.L.ao_get_light(self.h, _p(d["light"]))
        d["camera"] = np.zeros(1, camera_dt); self.L.ao_get_camera(self.h, _p(d["camera"]))
        d["pos"] = np.zeros((nv, 3), np.float32); d["nrm"] = np.zeros((nv, 3), np.float32)
        d["tris"] = np.zeros((nt, 3), np.int32); d["bbox"] = np.zeros(6, np.float32)
        self.L.ao_get_mesh(self.h, _p(d["pos"]), _p(d["nrm"]), _p(d["tris"]), _p(d["bbox"]))
        return d

    def set_spheres(self, a): a = np.ascontiguousarray(a, sphere_dt); self.L.ao_set_spheres(self.h, _p(a), len(a))
    def set_cornell(self, a): a = np.ascontiguousarray(a, cornell_dt); self.L.ao_set_cornell(self.h, _p(a))
    def set_flatlight(self, a): a = np.ascontiguousarray(a, flatlight_dt); self.L.ao_set_flatlight(self.h, _p(a))
    def set_materials(self, a): a = np.ascontiguousarray(a, material_dt); self.L.ao_set_materials(self.h, _p(a), len(a))
    def set_light(self, a): a = np.ascontiguousarray(a, light_dt); self.L.ao_set_light(self.h, _p(a))
    def set_camera(self, a): a = np.ascontiguousarray(a, camera_dt); self.L.ao_set_camera(self.h, _p(a))

    def set_mesh(self, pos, nrm, tris):
        pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        self.L.ao_set_mesh(self.h, _p(pos), _p(nrm), len(pos), _p(tris), len(tris))

    def set_mesh_tri_materials(self, mats):
        mats = np.ascontiguousarray(mats, np.int32)
        self.L.ao_set_mesh_tri_materials(self.h, _p(mats), len(mats))

    # --- queries ---------------------------------------------------------------------------------
    def light_sample_batch(self, u, p):
        u = np.ascontiguousarray(u, np.float32); p = np.ascontiguousarray(p, np.float32); n = len(p)
        pos = np.empty((n, 3), np.float32); pdf = np.empty(n, np.float32)
        self.L.ao_light_sample_batch(self.h, n, _p(u), _p(p), _p(pos), _p(pdf))
        return pos, pdf

    def light_eval_pdf_batch(self, p, d, t):
        p = np.ascontiguousarray(p, np.float32); d = np.ascontiguousarray(d, np.float32); t = np.ascontiguousarray(t, np.float32)
        out = np.empty(len(p), np.float32)
        self.L.ao_light_eval_pdf_batch(self.h, len(p), _p(p), _p(d), _p(t), _p(out))
        return out

    def mat_sample_batch(self, mat_id, r, ray_dir, normal):
        r = np.ascontiguousarray(r, np.float32); ray_dir = np.ascontiguousarray(ray_dir, np.float32)
        normal = np.ascontiguousarray(normal, np.float32); n = len(r)
        color = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32); pdf = np.empty(n, np.float32)
        spec = np.empty(n, np.int32)
        self.L.ao_mat_sample_batch(self.h, mat_id, n, _p(r), _p(ray_dir), _p(normal), _p(color), _p(d), _p(pdf), _p(spec))
        return color, d, pdf, spec

    def mat_eval_batch(self, mat_id, l, v, n_):
        l = np.ascontiguousarray(l, np.float32); v = np.ascontiguousarray(v, np.float32); n_ = np.ascontiguousarray(n_, np.float32)
        bxdf = np.empty((len(l), 3), np.float32); pdf = np.empty(len(l), np.float32)
        self.L.ao_mat_eval_batch(self.h, mat_id, len(l), _p(l), _p(v), _p(n_), _p(bxdf), _p(pdf))
        return bxdf, pdf

    def closest_hit_batch(self, rays, mesh_mode=MESH_REFERENCE_BF):
        rays = np.ascontiguousarray(rays, ray_dt)
        hits = np.zeros(len(rays), hit_dt)
        self.L.ao_closest_hit_batch(self.h, mesh_mode, len(rays), _p(rays), _p(hits))
        return hits

    def eye_ray(self, w, h, x, y, sub=-1):
        r = np.zeros(1, ray_dt)
        self.L.ao_eye_ray(self.h, w, h, x, y, sub, _p(r))
        return r[0]

    def primary_hits(self, w, h, mesh_mode=MESH_REFERENCE_BF):
        p = make_params(w, h, RT_DEBUG, mesh_mode=mesh_mode)
        hits = np.zeros(w * h, hit_dt)
        self.L.ao_primary_hits(self.h, C.byref(p), _p(hits))
        return hits.reshape(h, w)

    def compute_shadow(self, hit_pos, lpos, mesh_mode=MESH_REFERENCE_BF):
        a = np.asarray(hit_pos, np.float32); b = np.asarray(lpos, np.float32)
        return bool(self.L.ao_compute_shadow(self.h, mesh_mode, _p(a), _p(b)))

    def render(self, params, accum=None):
        """Returns (accum[H,W,3] float32, rays)."""
        if accum is None:
            accum = np.zeros((params.height, params.width, 3), np.float32)
        rays = self.L.ao_render(self.h, C.byref(params), _p(accum))
        return accum, int(rays)


def render_task_lock(scene, params, tasks):
    """the reference's tasking scheme (T tasks x full image, one global lock): (accum, rays); timing leg only"""
    accum = np.zeros((params.height, params.width, 3), np.float32)
    rays = lib().ao_render_task_lock(scene.h, C.byref(params), int(tasks), _p(accum))
    return accum, int(rays)


def resolve(accum, spp):
    accum = np.ascontiguousarray(accum, np.float32)
    n = accum.size // 3
    out = np.zeros(n, np.uint32)
    lib().ao_resolve(_p(accum), n, spp, _p(out))
    return out.reshape(accum.shape[:-1])


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().ao_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().ao_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def sincos_batch(x):
    """The repository's fixed binary32 sin/cos sequence (stands in for GNAT's run-time sin/cos) on an array."""
    x = np.ascontiguousarray(x, np.float32); s = np.empty_like(x); c = np.empty_like(x)
    lib().ao_sincos_batch(x.size, _p(x), _p(s), _p(c))
    return s, c


def expon_batch(left, right):
    """Left ** Right (stands in for GNAT's run-time "**") on arrays."""
    left = np.ascontiguousarray(left, np.float32); right = np.ascontiguousarray(np.broadcast_to(right, left.shape), np.float32)
    out = np.empty_like(left)
    lib().ao_expon_batch(left.size, _p(left), _p(right), _p(out))
    return out


def ada_pow(x, y):
    return lib().ao_pow(float(x), float(y))
