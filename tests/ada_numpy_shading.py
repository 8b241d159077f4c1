"""A SECOND, independent restatement of the SHADING part of the reference in numpy binary32: Lights.Sample / EvalPDF
(lights.adb), the BxDFs of materials.adb and the sampling maps of vector_math.adb.

Test infrastructure only.  Written directly from the Ada sources (file:line cited per function), NOT from
oracle/art_oracle.cpp: tests/test_oracle_kat.py demands that the oracle agrees with it bit for bit on 10^5 random inputs
per function, which pins the part of the oracle that round 1 only tied down by closed-form radiometry at 2-5 %.

What is NOT independent: sin, cos and "**" are run-time library calls in the reference (GNAT's Ada.Numerics, not in
tree); this repository replaces them by fixed binary32 operation sequences, and the restatement calls those through
`sincos` / `expon` callbacks (the oracle's batched det_sincos / ada_expon) -- everything around them is restated.
Where the reference would raise (sqrt of a negative argument: Argument_Error kills the task), oracle and kernels clamp
the argument to zero; the same is done here at the marked places.

numpy float32 array arithmetic is one correctly rounded IEEE binary32 operation per +,-,*,/,sqrt (no FMA contraction).
"""
import numpy as np

F = np.float32
M_PI = F(3.14159265358979323846)       # vector_math.ads:19
INV_PI = F(0.318309886183790671538)    # vector_math.ads:20
INFINITY = F(3.4028235e38)             # vector_math.ads:17
EPS_DIV = F(1.0e-20)                   # materials.adb:14, lights.adb:39
EPS_COS = F(1.0e-6)                    # materials.adb:15


def _min(a, b):                        # generic_vector_math.adb:19-26
    return np.where(a < b, a, b)


def _max(a, b):                        # :28-35
    return np.where(a >= b, a, b)


def _clamp(x, a, b):                   # :60-63  min(max(x, a), b)
    return _min(_max(x, a), b)


def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _cross(a, b):                      # generic_vector_math.adb:115-122
    return np.stack([a[..., 1] * b[..., 2] - b[..., 1] * a[..., 2],
                     a[..., 2] * b[..., 0] - b[..., 2] * a[..., 0],
                     a[..., 0] * b[..., 1] - b[..., 0] * a[..., 1]], -1)


def _scale(k, a):                      # float * float3 and float3 * float: (k*a.x, k*a.y, k*a.z), :148-156
    return np.asarray(k, F)[..., None] * a


def _normalize(a):                     # vector_math.adb:64-72
    with np.errstate(all="ignore"):
        l_inv = F(1.0) / np.sqrt(_dot(a, a))
    return l_inv[..., None] * a


def _length(a):                        # :74-77
    return np.sqrt(_dot(a, a))


def _reflect(d, n):                    # :79-82   normalize((normal * dot(dir, normal) * (-2.0)) + dir)
    return _normalize(_scale(F(-2.0), _scale(_dot(d, n), n)) + d)


def _sqrt_clamped(x):                  # ADA-RAISES site: sqrt of a negative argument
    return np.sqrt(_max(x, F(0.0)))


def pow_ada(left, right, expon):       # function pow, vector_math.adb:24-47 (the raising branches return what oracle/kernels return)
    left = np.asarray(left, F); right = np.broadcast_to(np.asarray(right, F), left.shape)
    with np.errstate(all="ignore"):
        gen = expon(left, right)
    out = gen
    out = np.where(right == 1.0, left, out)
    out = np.where(left == 1.0, F(1.0), out)
    out = np.where(left == 0.0, np.where(right < 0.0, F(np.inf), F(0.0)), out)
    out = np.where(right == 0.0, F(1.0), out)
    out = np.where(left < 0.0, F(0.0), out)
    out = np.where((left == 0.0) & (right == 0.0), F(1.0), out)
    return out.astype(F)


# ---------------------------------------------------------------------------------------------- vector_math.adb sampling maps
def get_perpendicular(a):              # :175-200
    xp, yp, zp = np.abs(a[..., 0]), np.abs(a[..., 1]), np.abs(a[..., 2])
    e = F(1.0e-5)
    first = (xp <= yp + e) & (xp <= zp + e)
    second = (~first) & (yp < xp + e) & (yp <= zp + e)
    lp = np.zeros_like(a)
    lp[..., 0] = np.where(first, F(1.0), F(0.0))
    lp[..., 1] = np.where(second, F(1.0), F(0.0))
    lp[..., 2] = np.where(~first & ~second, F(1.0), F(0.0))
    return _normalize(_cross(a, lp))


def _map_tail(deviation, direction, normal):     # :218-250 and, identically, :278-310
    ny = direction
    nx = get_perpendicular(ny)
    nz = _normalize(_cross(nx, ny))
    ny, nz = nz, ny                                                            # swap(ny, nz)
    res = _scale(deviation[..., 0], nx) + _scale(deviation[..., 1], ny) + _scale(deviation[..., 2], nz)
    inv_sign = np.where(_dot(direction, normal) >= 0.0, F(1.0), F(-1.0))
    under = inv_sign * _dot(res, normal) < 0.0
    with np.errstate(all="ignore"):
        nx2 = _normalize(_cross(normal, direction))
        nz2 = _normalize(_cross(nx2, ny))
        nz2 = np.where((_dot(nz2, res) < 0.0)[..., None], _scale(F(-1.0), nz2), nz2)
        res2 = _reflect(_scale(F(-1.0), res), nz2)
        res2 = np.where((_dot(res2, normal) < 0.0)[..., None], direction, res2)
    return np.where(under[..., None], res2, res)


def map_sample_to_cosine_dist(r1, r2, direction, normal, power, sincos, expon):      # :202-252
    sin_phi, cos_phi = sincos(F(2.0) * r1 * M_PI)
    cos_theta = expon(F(1.0) - r2, np.full_like(r2, F(1.0) / (F(power) + F(1.0))))
    sin_theta = _sqrt_clamped(F(1.0) - cos_theta * cos_theta)
    deviation = np.stack([sin_theta * cos_phi, sin_theta * sin_phi, cos_theta], -1)
    return _map_tail(deviation, direction, normal)


def map_sample_to_cosine_dist_fixed(r1, r2, direction, normal, power, sincos, expon):  # :255-312
    h = _sqrt_clamped(F(1.0) - expon(r1, np.full_like(r1, F(2.0) / (F(power) + F(1.0)))))
    s, c = sincos(F(2.0) * M_PI * r2)
    deviation = np.stack([h * c, h * s, expon(r1, np.full_like(r1, F(1.0) / (F(power) + F(1.0))))], -1)
    return _map_tail(deviation, direction, normal)


# ---------------------------------------------------------------------------------------------- lights.adb
def _distance_squared(a, b):           # :123-130
    diff = b - a
    return _dot(diff, diff)


def _coordinate_system(v1):            # :107-121
    with np.errstate(all="ignore"):
        big_x = np.abs(v1[..., 0]) > np.abs(v1[..., 1])
        inv_a = F(1.0) / np.sqrt(v1[..., 0] * v1[..., 0] + v1[..., 2] * v1[..., 2])
        inv_b = F(1.0) / np.sqrt(v1[..., 1] * v1[..., 1] + v1[..., 2] * v1[..., 2])
        zero = np.zeros_like(inv_a)
        v2a = np.stack([-v1[..., 2] * inv_a, zero, v1[..., 0] * inv_a], -1)
        v2b = np.stack([zero, v1[..., 2] * inv_b, -v1[..., 1] * inv_b], -1)
    v2 = np.where(big_x[..., None], v2a, v2b)
    return v2, _cross(v1, v2)


def sphere_light_eval_pdf(center, radius, area, p):                              # :242-255
    with np.errstate(all="ignore"):
        d2 = _distance_squared(p, center)
        inside = d2 - radius * radius < F(1.0e-4)
        sin2 = radius * radius / d2
        cos_max = np.sqrt(_max(F(0.0), F(1.0) - sin2))
        cone = F(1.0) / (F(2.0) * M_PI * (F(1.0) - cos_max))                    # UniformConePdf :153-159
    return np.where(inside, F(1.0) / area, cone).astype(F)


def sphere_light_sample(center, radius, area, u1, u2, p, sincos):                # :197-240 -> (pos, pdf)
    center = np.broadcast_to(np.asarray(center, F), p.shape)
    with np.errstate(all="ignore"):
        d2 = _distance_squared(p, center)
        inside = d2 - radius * radius < F(1.0e-4)
        # inside: UniformSampleSphere :133-142; res.pdf keeps its record default 1.0 (lights.ads:19)
        z = F(1.0) - F(2.0) * u1
        r = np.sqrt(_max(F(0.0), F(1.0) - z * z))
        s_in, c_in = sincos(F(2.0) * M_PI * u2)
        pos_in = center + _scale(radius, np.stack([r * c_in, r * s_in, z], -1))
        # outside: cone towards the centre
        wc = _normalize(center - p)
        wc_x, wc_y = _coordinate_system(wc)
        sin2 = radius * radius / d2
        cos_max = np.sqrt(_max(F(0.0), F(1.0) - sin2))
        costheta = (F(1.0) - u1) * cos_max + u1 * F(1.0)                          # lerp(u1, costhetamax, 1.0), vector_math.adb:59-62
        sintheta = _sqrt_clamped(F(1.0) - costheta * costheta)
        s, c = sincos(u2 * F(2.0) * M_PI)
        rdir = _scale(c * sintheta, wc_x) + _scale(s * sintheta, wc_y) + _scale(costheta, wc)   # UniformSampleCone :144-151
        rpos = p + _scale(F(1.0e-3), rdir)
        k = rpos - center                                                        # RaySphereIntersect :161-194
        b = _dot(k, rdir)
        cc = _dot(k, k) - radius * radius
        dd = b * b - cc
        sq = np.sqrt(_max(dd, F(0.0)))
        t1, t2 = -b - sq, -b + sq
        hx = np.where(dd >= 0.0, _min(t1, t2), -INFINITY)
        thit = np.where(hx < 0.0, _dot(center - p, _normalize(rdir)), hx)
        pos_out = rpos + _scale(thit, rdir)
        pdf_out = sphere_light_eval_pdf(center, radius, area, p)
    pos = np.where(inside[..., None], pos_in, pos_out).astype(F)
    pdf = np.where(inside, F(1.0), pdf_out).astype(F)
    return pos, pdf


def area_light_sample(box_min, box_max, normal, area, u1, u2, p):                # :56-79 -> (pos, pdf)
    pos = np.stack([box_min[0] + u1 * (box_max[0] - box_min[0]), np.full_like(u1, box_min[1]),
                    box_min[2] + u2 * (box_max[2] - box_min[2])], -1).astype(F)
    ray_dir = pos - p
    d = _length(ray_dir)
    with np.errstate(all="ignore"):
        ray_dir = _scale(F(1.0) / d, ray_dir)
        cos_theta = _max(_dot(ray_dir, _scale(F(-1.0), np.broadcast_to(np.asarray(normal, F), p.shape))), F(0.0))
        pdf = (F(1.0) / area) * d * d / _max(cos_theta, EPS_DIV)                 # PdfAtoW :42-45
    return pos, pdf.astype(F)


# ---------------------------------------------------------------------------------------------- materials.adb
def lambert_sample(kd, r1, r2, normal, sincos, expon):                            # :197-215
    new_dir = map_sample_to_cosine_dist(r1, r2, normal, normal, 1.0, sincos, expon)   # RandomCosineVectorOf, vector_math.adb:314-319
    cos_theta = _dot(new_dir, normal)
    pdf = np.abs(cos_theta) * INV_PI
    color = np.broadcast_to(_scale(INV_PI, np.asarray(kd, F)), new_dir.shape)
    color = np.where((cos_theta < EPS_COS)[..., None], F(0.0), color)
    return color.astype(F), new_dir, pdf.astype(F)


def lambert_eval(kd, l, n):                                                       # :217-226
    return np.broadcast_to(_scale(INV_PI, np.asarray(kd, F)), l.shape).astype(F), (_max(_dot(n, l), F(0.0)) * INV_PI).astype(F)


def mirror_sample(reflection, ray_dir, normal):                                   # :247-254
    next_dir = _reflect(ray_dir, normal)
    with np.errstate(all="ignore"):
        div = F(1.0) / _max(_dot(next_dir, normal), EPS_DIV)
    return _scale(div, np.broadcast_to(np.asarray(reflection, F), ray_dir.shape)).astype(F), next_dir, np.ones_like(div)


def _fresnel(cos1, eta_ext, eta_int):                                             # :79-99 with fresnelDielectric :70-76
    swap = cos1 < 0.0
    e_ext = np.where(swap, eta_int, eta_ext).astype(F); e_int = np.where(swap, eta_ext, eta_int).astype(F)
    with np.errstate(all="ignore"):
        sin2 = (e_ext / e_int) * np.sqrt(_max(F(0.0), F(1.0) - cos1 * cos1))
        cos2 = np.sqrt(_max(F(0.0), F(1.0) - sin2 * sin2))
        c1 = np.abs(cos1)
        a_ext, a_int = e_int, e_ext                                              # fresnelDielectric(abs(cosTheta1), cosTheta2, etaInt, etaExt)
        rs = (a_ext * c1 - a_int * cos2) / (a_ext * c1 + a_int * cos2)
        rp = (a_int * c1 - a_ext * cos2) / (a_int * c1 + a_ext * cos2)
        f = (rs * rs + rp * rp) / F(2.0)
    return np.where(sin2 > 1.0, F(1.0), f).astype(F)


def glass_sample(reflection, transparency, ior, ksi, ray_dir, normal):            # :285-331
    ior = F(ior)
    f = _fresnel(_dot(ray_dir, normal), np.full(ray_dir.shape[:-1], ior, F), np.ones(ray_dir.shape[:-1], F))     # ApplyFresnel :285-293
    refl = _scale(f, np.broadcast_to(np.asarray(reflection, F), ray_dir.shape))
    trans = _scale(F(1.0) - f, np.broadcast_to(np.asarray(transparency, F), ray_dir.shape))
    with np.errstate(all="ignore"):
        lr, lt = _length(refl), _length(trans)
        ksitrans = lt / (lr + lt)
        ksirefl = lr / (lr + lt)
        reflect_dir = _reflect(ray_dir, normal)
        # TotalInternalReflection :18-32 and refract :34-52
        cos_i = _dot(_scale(F(-1.0), ray_dir), normal)
        eta = np.where(cos_i < 0.0, F(1.0) / ior, ior).astype(F)
        k2 = F(1.0) - (F(1.0) - cos_i * cos_i) / (eta * eta)
        tir = k2 < 0.0
        flip = cos_i < 0.0
        cos_i2 = np.where(flip, -cos_i, cos_i)
        n2 = np.where(flip[..., None], _scale(F(-1.0), normal), normal)
        cos_t = np.sqrt(_max(k2, F(0.0)))
        wo = _scale(F(-1.0), ray_dir)
        wt = _normalize(_scale(F(1.0) / eta, _scale(F(-1.0), wo)) - _scale(cos_t - cos_i2 / eta, n2))
        take_refl = ksi > ksitrans
        next_dir = np.where(take_refl[..., None], reflect_dir, np.where(tir[..., None], reflect_dir, wt))
        bxdf = np.where(take_refl[..., None], _scale(F(1.0) / ksirefl, refl), _scale(F(1.0) / ksitrans, trans))
        div = F(1.0) / _max(np.abs(_dot(next_dir, normal)), EPS_DIV)
    return _scale(div, bxdf).astype(F), next_dir.astype(F), np.ones_like(div)


def phong_sample(reflection, cos_power, r1, r2, ray_dir, normal, sincos, expon):   # :363-387
    n = F(cos_power)
    r = _reflect(ray_dir, normal)
    next_dir = map_sample_to_cosine_dist_fixed(r1, r2, r, normal, n, sincos, expon)  # RandomCosineVectorOf, vector_math.adb:321-326
    cos_theta = _clamp(_dot(next_dir, r), F(0.0), M_PI * F(0.499995))
    pw = pow_ada(cos_theta, n, expon)
    refl = np.broadcast_to(np.asarray(reflection, F), ray_dir.shape)
    color = _scale(pw, _scale(INV_PI, _scale(F(0.5), _scale(n + F(2.0), refl))))  # mat.reflection*(cosPower + 2.0)*0.5*INV_PI*pow(...)
    pdf = pw * (n + F(1.0)) * (F(0.5) * INV_PI)
    cos_geo = _dot(next_dir, normal)
    with np.errstate(all="ignore"):
        div = F(1.0) / _max(np.abs(cos_geo), EPS_DIV)
    color = np.where((cos_geo < EPS_COS)[..., None], F(0.0), color)
    return _scale(div, color).astype(F), next_dir.astype(F), pdf.astype(F)


def phong_eval(reflection, cos_power, l, v, n_):                                   # :389-410
    n = F(cos_power)
    r = _reflect(_scale(F(-1.0), v), n_)
    cos_theta = _clamp(_dot(l, r), F(0.0), M_PI * F(0.499995))
    with np.errstate(all="ignore"):
        div = F(1.0) / _max(_dot(l, n_), EPS_DIV)
    return cos_theta, div, r, n, np.broadcast_to(np.asarray(reflection, F), l.shape)


def phong_eval_finish(parts, expon):
    cos_theta, div, r, n, refl = parts
    pw = pow_ada(cos_theta, n, expon)
    bxdf = _scale(div, _scale(pw, _scale(INV_PI, _scale(F(0.5), _scale(n + F(2.0), refl)))))
    pdf = pw * (n + F(1.0)) * (F(0.5) * INV_PI)
    return bxdf.astype(F), pdf.astype(F)
