"""A SECOND, independent restatement of the INTEGRATORS of the reference in numpy binary32: Integrator.DoPass and the three
PathTrace bodies of ray_tracer-integrators.adb (SimplePathTracer :82-126, PathTracerWithShadowRays :137-193,
PathTracerMIS :203-301), on top of the other two numpy restatements (tests/ada_numpy.py: rays, Find_Closest_Hit,
Compute_Shadow; tests/ada_numpy_shading.py: lights, BxDFs, sampling maps).

Test infrastructure only.  Written from the Ada sources, NOT from oracle/art_oracle.cpp; tests/test_oracle_kat.py demands
that a small frame rendered by the oracle equals this one bit for bit (every accumulator word and the ray count), for all three
integrators, with and without anti-aliasing.

The recursion is run level by level over all paths at once: what a level contributes (explicit light, throughput, or the
value a path ends with) is kept per level and the returns are folded innermost-first, which is the order the Ada recursion
evaluates `explicitColor + (abs(cosTheta)*bxdfVal)*PathTrace(...)` in.

Not independent (documented substitutions of this repository, the same on every side): the random stream (Philox4x32-10,
key = (seed, pixel), counter = (sample, bounce): light u1, u2, BSDF r1, r2) instead of Ada.Numerics.Float_Random, and the
fixed binary32 sequences for sin / cos / "**" handed in as `sincos` / `expon`.
"""
import numpy as np

import ada_numpy as an
import ada_numpy_shading as sh

F = np.float32
G_EPSILON = F(1.0e-5)                 # ray_tracer.ads:29
G_EPSILON_DIV = F(1.0e-20)            # ray_tracer.ads:30
MAT_NULL, MAT_LIGHT, MAT_LAMBERT, MAT_MIRROR, MAT_GLASS, MAT_PHONG = range(6)
PT_STUPID, PT_SHADOW, PT_MIS = 2, 3, 4


def philox4x32_10(ctr, key):
    """Salmon et al., SC'11 (Random123 constants); ctr [N,4], key [N,2] uint32 -> [N,4] uint32"""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0 = key[:, 0].astype(np.uint64); k1 = key[:, 1].astype(np.uint64)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]; p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + W0) & MASK; k1 = (k1 + W1) & MASK
    return np.stack(c, 1).astype(np.uint32)


def bounce_uniforms(seed, pixel, sample, bounce):
    """the four uniforms of one bounce: rnd_uniform(0.0, 1.0) = l + (h - l) * t (vector_math.adb:166-171) with t = (x >> 8) * 2^-24"""
    n = len(pixel)
    ctr = np.zeros((n, 4), np.uint32); ctr[:, 0] = sample; ctr[:, 1] = bounce
    key = np.zeros((n, 2), np.uint32); key[:, 0] = seed; key[:, 1] = pixel
    x = philox4x32_10(ctr, key)
    t = (x >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)
    return (F(0.0) + (F(1.0) - F(0.0)) * t).astype(F)


def _scale(k, a):
    return np.asarray(k, F)[..., None] * a


def _mat_sample(mats, mat_id, u, ray_dir, normal, sincos, expon):
    """SampleAndEvalBxDF dispatched on the material of each hit -> color, direction, pdf, pureSpecular"""
    n = len(mat_id)
    color = np.zeros((n, 3), F); direction = np.zeros((n, 3), F); pdf = np.ones(n, F); spec = np.zeros(n, bool)
    for m in np.unique(mat_id):
        sel = mat_id == m
        mt = mats[m]; kind = int(mt["type"])
        d, nr, r1, r2 = ray_dir[sel], normal[sel], u[sel, 2], u[sel, 3]
        if kind == MAT_LAMBERT:
            c, dd, p = sh.lambert_sample(mt["color"], r1, r2, nr, sincos, expon); s = False
        elif kind == MAT_MIRROR:
            c, dd, p = sh.mirror_sample(mt["color"], d, nr); s = True
        elif kind == MAT_GLASS:
            c, dd, p = sh.glass_sample(mt["color"], mt["color2"], mt["param"], r1, d, nr); s = True
        elif kind == MAT_PHONG:
            c, dd, p = sh.phong_sample(mt["color"], mt["param"], r1, r2, d, nr, sincos, expon); s = False
        else:                                                    # MaterialLight :163-166 -- never sampled: a path ends on it
            c, dd, p, s = np.zeros_like(d), np.zeros_like(d), np.ones(len(d), F), False
        color[sel] = c; direction[sel] = dd; pdf[sel] = p; spec[sel] = s
    return color, direction, pdf, spec


def _mat_eval(mats, mat_id, l, v, normal, expon):
    """EvalBxDF / EvalPDF dispatched on the material -> bxdf, pdf"""
    n = len(mat_id)
    bxdf = np.zeros((n, 3), F); pdf = np.ones(n, F)
    for m in np.unique(mat_id):
        sel = mat_id == m
        mt = mats[m]; kind = int(mt["type"])
        if kind == MAT_LAMBERT:
            b, p = sh.lambert_eval(mt["color"], l[sel], normal[sel])
        elif kind == MAT_PHONG:
            b, p = sh.phong_eval_finish(sh.phong_eval(mt["color"], mt["param"], l[sel], v[sel], normal[sel]), expon)
        else:                                                    # mirror, glass: (0,0,0) / 1.0 (materials.adb:256-264, 333-341)
            b, p = np.zeros((int(sel.sum()), 3), F), np.ones(int(sel.sum()), F)
        bxdf[sel] = b; pdf[sel] = p
    return bxdf, pdf


def path_trace(kind, o, d, pixel, sample, seed, max_depth, scene, mats, light, sincos, expon):
    """PathTrace(r, StartSample, Max_Trace_Depth) for rays o, d [N,3] -> (color [N,3], number of Find_Closest_Hit calls)"""
    n = len(o)
    o = o.astype(F).copy(); d = d.astype(F).copy()
    alive = np.ones(n, bool)
    prev_pdf = np.ones(n, F); prev_spec = np.ones(n, bool)       # StartSample = ((0,0,0), (0,0,0), 1.0, true), materials.ads:25
    is_light = np.array([int(m["type"]) == MAT_LIGHT for m in mats])
    emit = np.array([light["intensity"] if (int(m["type"]) == MAT_LIGHT and int(m["light"]) >= 0) else (0, 0, 0) for m in mats], F)
    center, radius, area, intensity = light["center"].astype(F), F(light["radius"]), F(light["surfaceArea"]), light["intensity"].astype(F)
    rect = int(light["shape"]) == 0                               # Light_Shape_Rect: AreaLight (lights.adb:47-90), else SphereLight

    def light_sample(u1, u2, p):                                  # Lights.Sample -> (pos, pdf)
        if rect:
            return sh.area_light_sample(light["boxMin"].astype(F), light["boxMax"].astype(F), light["normal"].astype(F), area, u1, u2, p)
        return sh.sphere_light_sample(center, radius, area, u1, u2, p, sincos)

    def light_eval_pdf(p, ray_dir, hit_dist):                     # Lights.EvalPDF
        if rect:                                                  # :81-85 PdfAtoW(AreaPDF, hitDist, max(dot(rayDir, -normal), 0))
            cos_t = an._max(an._dot(ray_dir, -np.broadcast_to(light["normal"].astype(F), ray_dir.shape)), F(0.0))
            return ((F(1.0) / area) * hit_dist * hit_dist / an._max(cos_t, sh.EPS_DIV)).astype(F)
        return sh.sphere_light_eval_pdf(center, radius, area, p)
    levels = []                                                  # per level: (bounced, ended, end_value, explicit, throughput)
    rays = 0
    for level in range(max_depth):                               # recursion_level = max_depth - level > 0
        idx = np.nonzero(alive)[0]
        if len(idx) == 0:
            break
        h = an.find_closest_hit(o[idx], d[idx], scene)
        rays += len(idx)
        ended = np.zeros(n, bool); end_value = np.zeros((n, 3), F)
        hit = h["is_hit"] == 1
        mat = h["mat"]
        lightm = hit & is_light[np.where(hit, mat, 0)]
        # ---- the path ends here: miss, or an emitter
        ended[idx[~hit | lightm]] = True
        if kind != PT_SHADOW and lightm.any():
            li = idx[lightm]
            facing = ~(an._dot(-d[li], h["normal"][lightm]) < F(0.0))                       # :103, :224
            e = emit[mat[lightm]]
            if kind == PT_MIS:                                                               # :229-242
                with np.errstate(all="ignore"):
                    lgt = light_eval_pdf(o[li], d[li], h["t"][lightm])
                    bs = prev_pdf[li]
                    w = np.where(prev_spec[li], F(1.0), bs * bs / (lgt * lgt + bs * bs)).astype(F)
                e = _scale(w, e)
            end_value[li] = np.where(facing[:, None], e, F(0.0))
        # ---- a surface: explicit light sample, then the next ray
        surf = hit & ~lightm
        si = idx[surf]
        explicit = np.zeros((n, 3), F); thr = np.zeros((n, 3), F); bounced = np.zeros(n, bool)
        if len(si):
            bounced[si] = True
            u = bounce_uniforms(seed, pixel[si], sample[si], level)
            nr = h["normal"][surf]; t = h["t"][surf]; mid = mat[surf]
            ro, rd = o[si], d[si]
            hpos = ro + rd * t[:, None]
            with np.errstate(all="ignore"):
                if kind != PT_STUPID:
                    lpos, lpdf = light_sample(u[:, 0], u[:, 1], hpos)
                    sdir = an._normalize(lpos - hpos)
                    in_shadow = an.compute_shadow(hpos, lpos, scene)
                    rays += len(si)
                    bxdf, bpdf = _mat_eval(mats, mid, sdir, -rd, nr, expon)
                    cos1 = an._max(an._dot(sdir, nr), F(0.0))
                    inten = np.broadcast_to(intensity, hpos.shape)
                    if kind == PT_SHADOW:                                                    # :175
                        ex = _scale(F(1.0) / an._max(lpdf, G_EPSILON_DIV), inten * _scale(cos1, bxdf))
                    else:                                                                    # :273-279
                        misw = lpdf * lpdf / (lpdf * lpdf + bpdf * bpdf)
                        ex = _scale(misw, _scale(F(1.0) / an._max(lpdf, G_EPSILON_DIV), inten) * _scale(cos1, bxdf))
                    explicit[si] = np.where(in_shadow[:, None], F(0.0), ex)
                color, ndir, pdf, spec = _mat_sample(mats, mid, u, rd, nr, sincos, expon)     # :116, :183, :291
                bxdf_val = _scale(F(1.0) / an._max(pdf, G_EPSILON_DIV), color)
                cos_t = an._dot(ndir, nr)
                thr[si] = _scale(np.abs(cos_t), bxdf_val)
                sign = np.where(cos_t >= F(0.0), F(1.0), F(-1.0)).astype(F)                   # vector_math.adb:49-56
                o[si] = hpos + _scale(G_EPSILON, _scale(sign, nr))                           # :120-122
                d[si] = ndir
            prev_pdf[si] = pdf; prev_spec[si] = spec
        alive = bounced
        levels.append((bounced, ended, end_value, explicit, thr))
    # ---- the returns, innermost first; a path that used all its levels gets (0,0,0) from recursion_level = 0
    res = np.zeros((n, 3), F)
    with np.errstate(all="ignore"):
        for bounced, ended, end_value, explicit, thr in reversed(levels):
            if kind == PT_STUPID:
                val = thr * res                                                              # abs(cosTheta)*bxdfVal*PathTrace(...), :124
            else:
                val = explicit + thr * res                                                   # :191, :299
            res = np.where(ended[:, None], end_value, np.where(bounced[:, None], val, res)).astype(F)
    return res, rays


def do_pass(kind, width, height, passes, anti_aliasing, max_depth, seed, scene, mats, light, camera, sincos, expon, first_sample=0,
            background=(0.0, 0.0, 0.0)):
    """Integrator.DoPass (:25-71) `passes` times -> (accum [H,W,3] float32 starting from zeros, rays)"""
    acc = np.zeros((height, width, 3), F)
    ys, xs = np.mgrid[0:height, 0:width]
    pixel = (ys * width + xs).reshape(-1).astype(np.uint32)
    rays = 0
    for p in range(passes):
        if anti_aliasing:
            color = np.broadcast_to(np.asarray(background, F), (height * width, 3)).astype(F)
            for i in range(4):
                o, d = an.eye_rays(width, height, camera["pos"], camera["matrix"], i)
                sample = np.full(height * width, first_sample + p * 4 + i, np.uint32)
                c, r = path_trace(kind, o.reshape(-1, 3), d.reshape(-1, 3), pixel, sample, seed, max_depth, scene, mats, light, sincos, expon)
                color = (color + c).astype(F); rays += r
        else:
            o, d = an.eye_rays(width, height, camera["pos"], camera["matrix"], -1)
            sample = np.full(height * width, first_sample + p, np.uint32)
            color, r = path_trace(kind, o.reshape(-1, 3), d.reshape(-1, 3), pixel, sample, seed, max_depth, scene, mats, light, sincos, expon)
            rays += r
        acc = (color.reshape(height, width, 3) + acc).astype(F)                              # colBuff(x,y) := color + colBuff(x,y)
    return acc, rays
