"""Synthetic scene data for BASELINE config C5 (not in the reference; pinned in SURVEY.md 8d):

  Cornell box, sphere light and camera of scene.adb; mesh = cube-sphere of 6 faces x n x n quads x 2 triangles
  (n = 913 -> 10 002 828 triangles), radius 1.2 about (0, 1.6, 2.5), radial displacement
  0.05 * sum_{k=0..3} 2^-k * sin(2^k * 7 * theta_k + phi_k) with phases phi_k from Philox4x32-10(key = (0x5EED0005, 0),
  counter = (k, 0, 0, 0)), theta_0 = longitude, theta_1 = latitude, theta_2 = longitude + latitude,
  theta_3 = longitude - latitude; CCW-outward winding, per-face unshared vertices, material id 2; plus 64 analytic
  spheres, centres uniform in the box (same key, counter = (1000 + i, 0, 0, 0)), radius 0.05 .. 0.15, materials cycling
  {1, 5, 8, 0}.

Host-side data generation only (numpy); nothing here is on the render path.
"""
import numpy as np

SEED_C5 = 0x5EED0005


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10 (Random123 constants).  ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32"""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k = np.array(key, dtype=np.uint64).reshape(-1, 2).copy()
    if len(k) == 1 and len(c) > 1:
        k = np.repeat(k, len(c), 0)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for r in range(10):
        if r > 0:
            k[:, 0] = (k[:, 0] + W0) & MASK
            k[:, 1] = (k[:, 1] + W1) & MASK
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        n0 = ((p1 >> np.uint64(32)) ^ c[:, 1] ^ k[:, 0]) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c[:, 3] ^ k[:, 1]) & MASK
        n3 = p0 & MASK
        c = np.stack([n0, n1, n2, n3], 1)
    return c.astype(np.uint32)


def _uniform(words):
    return (words >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def cube_sphere(n=913, radius=1.2, center=(0.0, 1.6, 2.5), seed=SEED_C5):
    """-> (pos[nv,3] f32, nrm[nv,3] f32, tris[nt,3] i32); nt = 12 n^2, nv = 6 (n+1)^2"""
    phases = _uniform(philox4x32_10(np.array([[k, 0, 0, 0] for k in range(4)], np.uint32), np.array([[seed, 0]], np.uint32))[:, 0])
    phases = phases.astype(np.float64) * 2.0 * np.pi
    u = np.linspace(-1.0, 1.0, n + 1)
    uu, vv = np.meshgrid(u, u, indexing="xy")
    ones = np.ones_like(uu)
    faces = [(ones, vv, -uu), (-ones, vv, uu), (uu, ones, -vv), (uu, -ones, vv), (uu, vv, ones), (-uu, vv, -ones)]
    pos_all, nrm_all, tri_all = [], [], []
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (ii * (n + 1) + jj).ravel()
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    base_tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 0)
    for f, (x, y, z) in enumerate(faces):
        p = np.stack([x, y, z], -1).reshape(-1, 3)
        p /= np.linalg.norm(p, axis=1, keepdims=True)
        lon = np.arctan2(p[:, 2], p[:, 0])
        lat = np.arcsin(np.clip(p[:, 1], -1.0, 1.0))
        th = [lon, lat, lon + lat, lon - lat]
        disp = sum(0.05 * 2.0 ** -k * np.sin(2.0 ** k * 7.0 * th[k] + phases[k]) for k in range(4))
        r = radius + disp
        pos = p * r[:, None] + np.asarray(center)[None, :]
        pos_all.append(pos.astype(np.float32))
        nrm_all.append(p.astype(np.float32))
        # orientation: make the winding CCW seen from outside (geometric normal along +p)
        t = base_tris
        a, b, c = pos[t[0, 0]], pos[t[0, 1]], pos[t[0, 2]]
        if np.dot(np.cross(b - a, c - a), p[t[0, 0]]) < 0:
            t = t[:, [0, 2, 1]]
        tri_all.append((t + f * (n + 1) * (n + 1)).astype(np.int32))
    return np.concatenate(pos_all), np.concatenate(nrm_all), np.concatenate(tri_all)


def c5_spheres(count=64, seed=SEED_C5, box_min=(-2.5, 0.0, 0.0), box_max=(2.5, 5.0, 5.0)):
    """64 analytic spheres -> list of (pos3, r, mat)"""
    ctr = np.array([[1000 + i, 0, 0, 0] for i in range(count)], np.uint32)
    w = _uniform(philox4x32_10(ctr, np.array([[seed, 0]], np.uint32))).astype(np.float64)
    mats = [1, 5, 8, 0]
    out = []
    lo, hi = np.asarray(box_min), np.asarray(box_max)
    for i in range(count):
        r = 0.05 + 0.10 * w[i, 3]
        c = lo + r + w[i, :3] * (hi - lo - 2 * r)
        out.append((c.astype(np.float32), np.float32(r), mats[i % 4]))
    return out


def c5_analytic(count=64, seed=SEED_C5):
    """(spheres, cornell) as structured arrays with the layouts of include/gcore.h: the 64 Philox spheres followed by the
    sphere light of scene.adb:116-144 (centre (0, 4.5, 1), r 0.5, material 4), and the Cornell box of scene.ads:75-80"""
    sphere_dt = np.dtype([("pos", "<f4", 3), ("r", "<f4"), ("mat", "<i4")])
    cornell_dt = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("mat_indices", "<i4", 6), ("normals", "<f4", 18)])
    sp = np.zeros(count + 1, sphere_dt)
    for i, (c, r, m) in enumerate(c5_spheres(count, seed)):
        sp[i] = (c, r, m)
    sp[count] = ((0.0, 4.5, 1.0), 0.5, 4)
    box = np.zeros(1, cornell_dt)
    box["bmin"] = (-2.5, 0.0, 0.0); box["bmax"] = (2.5, 5.0, 5.0)
    box["mat_indices"] = (2, 3, 1, 1, 8, 1)
    box["normals"] = (1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1)
    return sp, box


def install_c5(scene, n=913, oracle_scene=None):
    """Configure `scene` (package Scene, already Init()-ed: materials, light, camera of scene.adb) as BASELINE config C5;
    optionally mirror it into an oracle scene (tests).  Returns (pos, nrm, tris)."""
    pos, nrm, tris = cube_sphere(n)
    sp, box = c5_analytic()
    scene.set_analytic(sp, box)
    scene.set_mesh(pos, nrm, tris, mesh_mode=1)          # MESH_NEAREST -> LBVH
    if oracle_scene is not None:
        oracle_scene.set_spheres(sp)
        oracle_scene.set_mesh(pos, nrm, tris)
    return pos, nrm, tris
