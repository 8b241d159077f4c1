"""A SECOND, independent restatement of the primary-ray part of the reference, vectorised in numpy binary32.

Test infrastructure only.  Written directly from the Ada sources (file:line cited per function), NOT from
oracle/art_oracle.cpp, so that the oracle (the C++ restatement every GPU parity test leans on) is itself pinned by an
independent derivation: tests/test_oracle_kat.py demands that both agree bit for bit on every field of the Hit record
for every primary ray of BASELINE config C2.

numpy float32 array arithmetic is one correctly rounded IEEE binary32 operation per +,-,*,/,sqrt (no FMA contraction),
which is the arithmetic GNAT -O0 emits on x86-64 (scalar SSE).
"""
import numpy as np

F = np.float32
INFINITY = F(3.4028235e38)            # vector_math.ads:17  infinity = Float'Last
EPS_DIV = F(1.0e-25)                  # geometry.adb:236


def _min(a, b):                       # generic_vector_math.adb:19-26   if a < b then a else b
    return np.where(a < b, a, b)


def _max(a, b):                       # generic_vector_math.adb:28-35   if a >= b then a else b
    return np.where(a >= b, a, b)


def _dot(a, b):                       # generic_vector_math.adb:110-113  a.x*b.x + a.y*b.y + a.z*b.z
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _cross(a, b):                     # generic_vector_math.adb:115-122
    return np.stack([a[..., 1] * b[..., 2] - b[..., 1] * a[..., 2],
                     a[..., 2] * b[..., 0] - b[..., 2] * a[..., 0],
                     a[..., 0] * b[..., 1] - b[..., 0] * a[..., 1]], -1)


def _normalize(a):                    # vector_math.adb:64-72   l_inv := 1.0/sqrt(dot(a,a)); (l_inv*a.x, ...)
    with np.errstate(all="ignore"):
        l_inv = F(1.0) / np.sqrt(_dot(a, a))
    return l_inv[..., None] * a


def eye_rays(width, height, cam_pos, cam_matrix, sub=-1):
    """ray_tracer.adb:61-97 (EyeRayDirection / Generate4RayDirections) + ray_tracer-integrators.adb:46,58
    (r.direction := normalize(cam.matrix*dir)); returns origins[H,W,3], directions[H,W,3]"""
    y, x = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    if sub < 0:
        offx = offy = F(0.5)
    else:                             # order (1/3,1/3) (1/3,2/3) (2/3,1/3) (2/3,2/3), ray_tracer.adb:76-90
        third, two_thirds = F(1.0) / F(3.0), F(2.0) / F(3.0)
        offx = two_thirds if (sub & 2) else third
        offy = two_thirds if (sub & 1) else third
    fov = F(np.pi) / F(2.0)
    tan_half = F(np.tan(np.float64(fov / F(2.0))))               # safe_tan(fov/2.0): tan(0.78539819) rounds to 1.0
    assert tan_half == F(1.0)
    rx = x.astype(F) + offx - (F(width) / F(2.0))
    ry = y.astype(F) + offy - (F(height) / F(2.0))
    rz = np.full_like(rx, -F(width) / tan_half)
    d = _normalize(np.stack([rx, ry, rz], -1))
    m = np.asarray(cam_matrix, F).reshape(4, 4)
    md = np.stack([m[r, 0] * d[..., 0] + m[r, 1] * d[..., 1] + m[r, 2] * d[..., 2] + m[r, 3] for r in range(3)], -1)   # vector_math.adb:137-144
    d = _normalize(md)
    o = np.broadcast_to(np.asarray(cam_pos, F), d.shape).copy()
    return o, d


def intersect_box(o, d, bmin, bmax):
    """geometry.adb:146-191 -> (is_hit, tmin, tmax)"""
    with np.errstate(all="ignore"):
        inv = F(1.0) / d
        lo = (bmax[0] - o[..., 0]) * inv[..., 0]
        hi = (bmin[0] - o[..., 0]) * inv[..., 0]
        lo1 = (bmax[1] - o[..., 1]) * inv[..., 1]
        hi1 = (bmin[1] - o[..., 1]) * inv[..., 1]
        lo2 = (bmax[2] - o[..., 2]) * inv[..., 2]
        hi2 = (bmin[2] - o[..., 2]) * inv[..., 2]
        tmin = _min(lo, hi)
        tmax = _max(lo, hi)
        tmin = _max(tmin, _min(lo1, hi1))
        tmax = _min(tmax, _max(lo1, hi1))
        tmin = _max(tmin, _min(lo2, hi2))
        tmax = _min(tmax, _max(lo2, hi2))
        return (tmax > 0) & (tmin <= tmax), tmin, tmax


def find_closest_hit(o, d, scene):
    """scene.adb:56-86 for rays o,d of shape [N,3]; scene = dict(spheres, cornell, materials, pos, nrm, tris, bbox)
    (sphere-light scene: IntersectFlatLight is skipped, scene.adb:65-67).  Returns a dict of arrays with the fields of
    Geometry.Hit (geometry.ads:57-67); `mat` is the index of the resolved material."""
    n = len(o)
    sph = scene["spheres"]
    # ---- hits(1): IntersectAllSpheres, geometry.adb:48-115
    min_t = np.full(n, INFINITY, F)
    min_i = np.zeros(n, np.int32)
    with np.errstate(all="ignore"):
        for i in range(len(sph)):
            c = np.asarray(sph["pos"][i], F)
            r = F(sph["r"][i])
            k = o - c
            b = _dot(k, d)
            cc = _dot(k, k) - r * r
            dd = b * b - cc
            ok = dd >= 0
            sqrtd = np.sqrt(np.where(ok, dd, F(0)))
            t1 = -b - sqrtd
            t2 = -b + sqrtd
            c1 = ok & (t1 > 0) & (t1 < min_t)
            c2 = ok & ~c1 & (t2 > 0) & (t2 < min_t)
            min_t = np.where(c1, t1, np.where(c2, t2, min_t))
            min_i = np.where(c1 | c2, i, min_i)
    s_hit = (min_t > 0) & (min_t < INFINITY)
    s_t = np.where(s_hit, min_t, F(1.0))                          # :85-87
    cpos = np.asarray(sph["pos"], F)[min_i]
    with np.errstate(all="ignore"):
        s_nrm = _normalize((o + d * min_t[:, None]) - cpos)       # :88
    s_nrm = np.where(s_hit[:, None], s_nrm, np.asarray([0, 1, 0], F))
    s_mat = np.asarray(sph["mat"], np.int32)[min_i]

    # ---- hits(2): IntersectCornellBox, geometry.adb:193-229
    box = scene["cornell"]
    bmin, bmax = np.asarray(box["bmin"], F).reshape(3), np.asarray(box["bmax"], F).reshape(3)
    b_in, _, b_tmax = intersect_box(o, d, bmin, bmax)
    with np.errstate(all="ignore"):
        p = o + b_tmax[:, None] * d
        eps = F(1.0e-5)
        plane = np.zeros(n, np.int32)
        plane = np.where(np.abs(p[:, 0] - bmin[0]) < eps, 0, plane)
        plane = np.where(np.abs(p[:, 0] - bmax[0]) < eps, 1, plane)
        plane = np.where(np.abs(p[:, 1] - bmin[1]) < eps, 2, plane)
        plane = np.where(np.abs(p[:, 1] - bmax[1]) < eps, 3, plane)
        plane = np.where(np.abs(p[:, 2] - bmin[2]) < eps, 4, plane)
        plane = np.where(np.abs(p[:, 2] - bmax[2]) < eps, 5, plane)
    b_hit = b_in & (plane != 5)
    b_t = b_tmax
    b_matid = np.asarray(box["mat_indices"], np.int32).reshape(6)[plane]
    b_nrm = np.asarray(box["normals"], F).reshape(6, 3)[plane]

    # ---- hits(3): IntersectFlatLight, geometry.adb:118-143 -- only in the rect-light variant of the scene (scene.adb:65-67)
    fl = scene.get("flatlight")
    if fl is not None:
        fmin, fmax = np.asarray(fl["boxMin"], F).reshape(3), np.asarray(fl["boxMax"], F).reshape(3)
        with np.errstate(all="ignore"):
            inv_dir_y = F(1.0) / d[:, 1]
            f_t = (fmax[1] - o[:, 1]) * inv_dir_y
            hp = o + f_t[:, None] * d
            f_hit = (hp[:, 0] > fmin[0]) & (hp[:, 0] < fmax[0]) & (hp[:, 2] > fmin[2]) & (hp[:, 2] < fmax[2]) & (f_t >= 0)
    else:
        f_hit, f_t = None, None

    # ---- hits(4): IntersectMeshBF, geometry.adb:266-323 (+ IntersectTriangle :231-263)
    pos, nrm, tris, bbox = scene["pos"], scene["nrm"], scene["tris"], scene["bbox"]
    m_in, _, _ = intersect_box(o, d, np.asarray(bbox[:3], F), np.asarray(bbox[3:], F))
    w_min = np.zeros(n, F)
    w_max = np.full(n, F(1000000.0), F)
    m_hit = np.zeros(n, bool)
    m_u = np.zeros(n, F); m_v = np.zeros(n, F)
    m_tri = np.zeros(n, np.int32)
    with np.errstate(all="ignore"):
        for i in range(len(tris)):
            A, B, C = pos[tris[i, 0]], pos[tris[i, 1]], pos[tris[i, 2]]
            edge1 = (B - A).astype(F); edge2 = (C - A).astype(F)
            pvec = _cross(d, np.broadcast_to(edge2, d.shape))
            tvec = o - A
            qvec = _cross(tvec, np.broadcast_to(edge1, d.shape))
            inv_det = F(1.0) / _max(_dot(np.broadcast_to(edge1, d.shape), pvec), EPS_DIV)
            v = _dot(tvec, pvec) * inv_det
            u = _dot(qvec, d) * inv_det
            t = _dot(np.broadcast_to(edge2, d.shape), qvec) * inv_det
            ok = m_in & (v > 0) & (u > 0) & (u + v < F(1.0)) & (t > w_min) & (t < w_max)
            w_min = np.where(ok, t, w_min)                        # nearestHit := tmpHit (tmin = t, tmax = t + 1e-6)
            w_max = np.where(ok, t + F(1.0e-6), w_max)
            m_u = np.where(ok, u, m_u); m_v = np.where(ok, v, m_v)
            m_tri = np.where(ok, i, m_tri)
            m_hit = m_hit | ok
    tri = tris[m_tri]
    w = F(1.0) - m_u - m_v
    m_nrm = w[:, None] * nrm[tri[:, 0]] + m_v[:, None] * nrm[tri[:, 1]] + m_u[:, None] * nrm[tri[:, 2]]     # :301
    m_tx = w * F(0) + m_v * F(0) + m_u * F(0)                     # uvs are zeroed by the loader, geometry.adb:565-566
    m_t = w_min                                                   # t => nearestHit.tmin

    # ---- nearest of the candidates, strict '<' in order (scene.adb:71-78)
    nearest = np.zeros(n, np.int32)                               # index 0 = hits(1)
    dist = np.full(n, INFINITY, F)
    for idx, (hit, t) in enumerate([(s_hit, s_t), (b_hit, b_t), (f_hit, f_t), (m_hit, m_t)]):
        if hit is None:
            continue
        take = hit & (t < dist)
        nearest = np.where(take, idx, nearest)
        dist = np.where(take, t, dist)

    out = {}
    if f_hit is None:
        f_hit, f_t = np.zeros(n, bool), np.zeros(n, F)
    sel = lambda a, b, c, q=None: np.where(nearest == 0, a, np.where(nearest == 1, b, np.where(nearest == 2, c if q is None else q, c)))
    out["is_hit"] = sel(s_hit, b_hit, m_hit, f_hit).astype(np.int32)
    out["prim_type"] = sel(1, 0, 2, 3).astype(np.int32)           # geometry.ads: Plane 0, Sphere 1, Triangle 2, Quad 3
    out["prim_index"] = sel(min_i, plane, m_tri, 0).astype(np.int32)
    out["matId"] = sel(0, b_matid, 2, 0).astype(np.int32)
    out["mat"] = sel(s_mat, b_matid, 2, 4).astype(np.int32)       # scene.adb:80-82; the flat light carries materials(4) (scene.adb:66)
    out["t"] = sel(s_t, b_t, m_t, f_t).astype(F)
    out["normal"] = np.where((nearest == 0)[:, None], s_nrm, np.where((nearest == 1)[:, None], b_nrm,
                             np.where((nearest == 2)[:, None], np.asarray([0, -1, 0], F), m_nrm))).astype(F)
    out["tx"] = sel(F(0), F(0), m_tx, F(0)).astype(F)
    out["ty"] = out["tx"].copy()
    return out


def _max3(a, b, c):                  # generic_vector_math.adb:48-57
    return np.where((a >= b) & (a >= c), a, np.where((b >= c) & (b >= a), b, c))


def compute_shadow(hit_pos, lpos, scene):
    """ray_tracer.adb:100-132 for arrays [N,3] -> bool[N] (in_shadow).  Uses the CLOSEST hit and two ad-hoc epsilons; the
    emitter sphere itself is geometry and is hit at t ~= maxDist, so in binary32 a sizeable share of unoccluded samples
    comes out "in shadow" -- reference behaviour that the oracle and the GPU path reproduce bit for bit."""
    hit_pos = np.asarray(hit_pos, F); lpos = np.asarray(lpos, F)
    eps = _max3(np.abs(hit_pos[:, 0]), np.abs(hit_pos[:, 1]), np.abs(hit_pos[:, 2])) * F(0.000000001)
    eps = _max(eps, F(1.0e-30))
    sdir = _normalize(lpos - hit_pos)
    so = hit_pos + sdir * eps[:, None]
    h = find_closest_hit(so, sdir, scene)
    diff = hit_pos - lpos
    max_dist = np.sqrt(_dot(diff, diff))
    eps2 = _max(max_dist * F(0.000001), F(1.0e-30))
    return (h["is_hit"] == 1) & (h["t"] < max_dist - eps2) & (h["t"] > F(10.0) * eps)


def to_u32_round(v):
    """Ada float -> Unsigned_32 conversion: round to nearest, ties away from zero (ray_tracer.adb:51-53)"""
    v = np.asarray(v, np.float64)
    return np.floor(v + 0.5).astype(np.uint32)


def resolve(accum, spp):
    """ray_tracer.adb:281-291 + ToneMapping :19-38 + ColorToUnsigned_32 :41-57"""
    norm_c = F(1.0) / F(spp)
    c = np.sqrt(norm_c * np.asarray(accum, F))                    # ** (1.0/g_gamma), g_gamma = 2.0
    c = np.where(c < F(1.0), c, F(1.0))                           # min(c, 1.0)
    q = to_u32_round(c * F(255.0))
    return q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16)
