// art_shade.cuh -- light sampling, BxDFs and the shadow test, fused into the render kernel.
//
// Replaces: lights.adb:42-266, materials.adb:18-410, vector_math.adb:175-326 (sampling maps),
// ray_tracer.adb:100-132 (Compute_Shadow).
#pragma once

#include "art_scene.cuh"

namespace art {

// ---------------------------------------------------------------- lights.adb
struct LightSample { f3 pos; f3 intensity; float pdf; };          // lights.ads:15-20 (dir is never read)

ART_D float pdf_a_to_w(float aPdfA, float aDist, float aCosThere) {                   // lights.adb:42-45
  return aPdfA * aDist * aDist / max2(aCosThere, 1.0e-20f);
}

ART_D float distance_squared(f3 a, f3 b) { f3 diff = b - a; return dot(diff, diff); } // :123-130

ART_D float light_eval_pdf(f3 lluminatingPoint, f3 rayDir, float hitDist) {
  if (g_scn.light_shape == LIGHT_RECT) {                                              // :81-85
    float cosTheta = max2(dot(rayDir, neg(ld3(g_scn.l_normal))), 0.0f);
    return pdf_a_to_w(1.0f / g_scn.l_area, hitDist, cosTheta);
  }
  f3 center = ld3(g_scn.l_center);                                                    // :242-255
  float radius = g_scn.l_radius;
  if (distance_squared(lluminatingPoint, center) - radius * radius < 1.0e-4f) return 1.0f / g_scn.l_area;
  float sinThetaMax2 = radius * radius / distance_squared(lluminatingPoint, center);
  float cosThetaMax  = sqrtf(max2(0.0f, 1.0f - sinThetaMax2));
  return 1.0f / (2.0f * ART_PI * (1.0f - cosThetaMax));                               // UniformConePdf :153-159
}

ART_D LightSample light_sample(float u1, float u2, f3 lluminatingPoint) {
  LightSample res;
  res.intensity = ld3(g_scn.l_intensity);
  res.pdf = 1.0f;
  if (g_scn.light_shape == LIGHT_RECT) {                                              // :56-79
    res.pos.x = g_scn.l_boxMin[0] + u1 * (g_scn.l_boxMax[0] - g_scn.l_boxMin[0]);
    res.pos.y = g_scn.l_boxMin[1];
    res.pos.z = g_scn.l_boxMin[2] + u2 * (g_scn.l_boxMax[2] - g_scn.l_boxMin[2]);
    f3 rayDir = res.pos - lluminatingPoint;
    float dlen = length(rayDir);
    rayDir = rayDir * (1.0f / dlen);
    float cosTheta = max2(dot(rayDir, neg(ld3(g_scn.l_normal))), 0.0f);
    res.pdf = pdf_a_to_w(1.0f / g_scn.l_area, dlen, cosTheta);
    return res;
  }
  f3 center = ld3(g_scn.l_center);                                                    // :197-240
  float radius = g_scn.l_radius;
  float dist2 = distance_squared(lluminatingPoint, center);
  if (dist2 - radius * radius < 1.0e-4f) {
    float z = 1.0f - 2.0f * u1;                                                       // UniformSampleSphere :133-142
    float r = sqrtf(max2(0.0f, 1.0f - z * z));
    float phi = 2.0f * ART_PI * u2;
    float s, c; det_sincos(phi, s, c);
    res.pos = center + radius * mk3(r * c, r * s, z);
    return res;                                                                       // pdf stays 1.0 (lights.ads:19)
  }
  f3 wc = normalize(center - lluminatingPoint);
  f3 wcX, wcY;                                                                        // CoordinateSystem :107-121
  if (fabsf(wc.x) > fabsf(wc.y)) {
    float invLen = 1.0f / sqrtf(wc.x * wc.x + wc.z * wc.z);
    wcX = mk3(-wc.z * invLen, 0.0f, wc.x * invLen);
  } else {
    float invLen = 1.0f / sqrtf(wc.y * wc.y + wc.z * wc.z);
    wcX = mk3(0.0f, wc.z * invLen, -wc.y * invLen);
  }
  wcY = cross(wc, wcX);
  float sinThetaMax2 = radius * radius / dist2;
  float cosThetaMax  = sqrtf(max2(0.0f, 1.0f - sinThetaMax2));
  float costheta = lerp_ada(u1, cosThetaMax, 1.0f);                                   // UniformSampleCone :144-151
  float sintheta = sqrt_clamped(1.0f - costheta * costheta);
  float phi = u2 * 2.0f * ART_PI;
  float s, c; det_sincos(phi, s, c);
  f3 rdir = c * sintheta * wcX + s * sintheta * wcY + costheta * wc;
  f3 rpos = lluminatingPoint + rdir * 1.0e-3f;
  // RaySphereIntersect :161-194
  f3 k = rpos - center;
  float b = dot(k, rdir);
  float cc = dot(k, k) - radius * radius;
  float dd = b * b - cc;
  float hx;
  if (dd >= 0.0f) { float sqrtd = sqrtf(dd); float t1 = -b - sqrtd, t2 = -b + sqrtd; hx = min2(t1, t2); }
  else hx = -ART_FLT_LAST;
  float thit;
  if (hx < 0.0f) thit = dot(center - lluminatingPoint, normalize(rdir));
  else thit = hx;
  res.pos = rpos + thit * rdir;
  res.pdf = 1.0f / (2.0f * ART_PI * (1.0f - cosThetaMax));                            // EvalPDF(l, p, rdir, thit) :237, same dist2 test as above
  return res;
}

// ---------------------------------------------------------------- vector_math.adb sampling maps
ART_D f3 get_perpendicular(f3 a) {                                                    // :175-200
  float xp = fabsf(a.x), yp = fabsf(a.y), zp = fabsf(a.z);
  f3 lp;
  if ((xp <= yp + 1.0e-5f) && (xp <= zp + 1.0e-5f)) lp = mk3(1.0f, 0.0f, 0.0f);
  else if ((yp < xp + 1.0e-5f) && (yp <= zp + 1.0e-5f)) lp = mk3(0.0f, 1.0f, 0.0f);
  else lp = mk3(0.0f, 0.0f, 1.0f);
  return normalize(cross(a, lp));
}

ART_D f3 map_tail(f3 deviation, f3 direction, f3 normal) {                            // :218-250 / :278-310
  f3 ny = direction;
  f3 nx = get_perpendicular(ny);
  f3 nz = normalize(cross(nx, ny));
  f3 tmp = ny; ny = nz; nz = tmp;
  f3 res = nx * deviation.x + ny * deviation.y + nz * deviation.z;
  float invSign = (dot(direction, normal) >= 0.0f) ? 1.0f : -1.0f;
  if (invSign * dot(res, normal) < 0.0f) {
    nx = normalize(cross(normal, direction));
    nz = normalize(cross(nx, ny));
    if (dot(nz, res) < 0.0f) nz = neg(nz);
    res = reflect(neg(res), nz);
    if (dot(res, normal) < 0.0f) res = direction;
  }
  return res;
}

// ---------------------------------------------------------------- materials.adb
struct MatSampleD { f3 color; f3 direction; float pdf; bool pureSpecular; };          // materials.ads:18-23

ART_D float fresnel_ada(float cosTheta1, float a_etaExt, float a_etaInt) {            // materials.adb:70-99
  float etaExt = a_etaExt, etaInt = a_etaInt;
  if (cosTheta1 < 0.0f) { float tmp = etaExt; etaExt = etaInt; etaInt = tmp; }
  float sinTheta2 = (etaExt / etaInt) * sqrtf(max2(0.0f, 1.0f - cosTheta1 * cosTheta1));
  if (sinTheta2 > 1.0f) return 1.0f;
  float cosTheta2 = sqrtf(max2(0.0f, 1.0f - sinTheta2 * sinTheta2));
  float c1 = fabsf(cosTheta1);
  // fresnelDielectric(c1, cosTheta2, etaExt => etaInt, etaInt => etaExt)
  float eE = etaInt, eI = etaExt;
  float Rs = (eE * c1 - eI * cosTheta2) / (eE * c1 + eI * cosTheta2);
  float Rp = (eI * c1 - eE * cosTheta2) / (eI * c1 + eE * cosTheta2);
  return (Rs * Rs + Rp * Rp) / 2.0f;
}

// "**" with a run-time exponent, kept out of line: one copy of the exp/log kernels serves every call site
__device__ __noinline__ float ada_expon_call(float left, float right) { return ada_expon(left, right); }

// function pow (vector_math.adb:24-47) through the out-of-line "**"
ART_D float pow_ada_call(float left, float right) {
  if (left == 0.0f && right == 0.0f) return 1.0f;
  else if (left < 0.0f) return 0.0f;
  else if (right == 0.0f) return 1.0f;
  else if (left == 0.0f) { return (right < 0.0f) ? __int_as_float(0x7f800000) : 0.0f; }
  else if (left == 1.0f) return 1.0f;
  else if (right == 1.0f) return left;
  else return ada_expon_call(left, right);
}

// SampleAndEvalBxDF of every material type (materials.adb:163-166, 197-215, 247-254, 295-331, 363-387) in one pass whose
// expensive common pieces are executed once, with the lanes of all types that need them converged:
//   refl = reflect(ray_dir, normal)                  mirror :251, glass :313/:322, Phong :371  (passed in by the caller)
//   sin/cos of the azimuth                           Lambert vector_math.adb:210-211, Phong :265-266
//   tangent frame + under-surface fix-up (map_tail)  Lambert :218-250, Phong :278-310
// Every value is produced by the same operation sequence as in the per-type functions of the reference.
ART_D MatSampleD mat_sample(const DMaterial& mat, float r1, float r2, f3 ray_dir, f3 normal, f3 refl) {
  MatSampleD ms;
  ms.color = mk3(0.0f, 0.0f, 0.0f); ms.direction = mk3(0.0f, 0.0f, 0.0f); ms.pdf = 1.0f; ms.pureSpecular = false;   // MaterialLight / null
  const int type = mat.type;
  const f3 refl_c = ld3(mat.color);
  const float param = mat.param;

  if (type == MAT_LAMBERT || type == MAT_PHONG) {
    const bool lam = (type == MAT_LAMBERT);
    // azimuth: Lambert sin/cos(2.0*r1*M_PI), Phong cos/sin(2.0*M_PI*r2)
    const float angle = lam ? (2.0f * r1 * ART_PI) : (2.0f * ART_PI * r2);
    float sn, cs;
    det_sincos(angle, sn, cs);
    f3 deviation;
    if (lam) {
      float cos_theta = ada_expon(1.0f - r2, 1.0f / (1.0f + 1.0f));          // (1.0-r2) ** (1.0/(e+1.0)), e = 1.0
      float sin_theta = sqrt_clamped(1.0f - cos_theta * cos_theta);
      deviation = mk3(sin_theta * cs, sin_theta * sn, cos_theta);
    } else {
      float hh = sqrt_clamped(1.0f - ada_expon_call(r1, 2.0f / (param + 1.0f)));
      deviation.x = hh * cs;
      deviation.y = hh * sn;
      deviation.z = ada_expon_call(r1, 1.0f / (param + 1.0f));
    }
    const f3 axis = lam ? normal : refl;
    const f3 nextDir = map_tail(deviation, axis, normal);
    if (lam) {                                                                // :204-213
      float cosTheta = dot(nextDir, normal);
      ms.pdf = fabsf(cosTheta) * ART_INV_PI;
      f3 color = refl_c * ART_INV_PI;
      if (cosTheta < 1.0e-6f) color = mk3(0.0f, 0.0f, 0.0f);
      ms.color = color; ms.direction = nextDir;
    } else {                                                                  // :374-385
      float cosTheta = clampf(dot(nextDir, refl), 0.0f, ART_PI * 0.499995f);
      float pw = pow_ada_call(cosTheta, param);
      f3 color = refl_c * (param + 2.0f) * 0.5f * ART_INV_PI * pw;
      ms.pdf = pw * (param + 1.0f) * (0.5f * ART_INV_PI);
      float cosThetaGeo = dot(nextDir, normal);
      float cosThetaDiv = 1.0f / max2(fabsf(cosThetaGeo), 1.0e-20f);
      if (cosThetaGeo < 1.0e-6f) color = mk3(0.0f, 0.0f, 0.0f);
      ms.color = color * cosThetaDiv; ms.direction = nextDir;
    }
  } else if (type == MAT_MIRROR) {                                            // :247-254
    float cosThetaDiv = 1.0f / max2(dot(refl, normal), 1.0e-20f);
    ms.color = refl_c * cosThetaDiv; ms.direction = refl; ms.pdf = 1.0f; ms.pureSpecular = true;
  } else if (type == MAT_GLASS) {                                             // :285-331
    f3 trans = ld3(mat.color2);
    float ior = param;
    float f = fresnel_ada(dot(ray_dir, normal), ior, 1.0f);
    f3 rf = f * refl_c;
    trans = (1.0f - f) * trans;
    float lr = length(rf), lt = length(trans);
    float ksitrans = lt / (lr + lt);
    float ksirefl  = lr / (lr + lt);
    f3 nextDirection = refl, bxdf;
    if (r1 > ksitrans) { bxdf = rf * (1.0f / ksirefl); }
    else {
      bxdf = trans * (1.0f / ksitrans);
      // TotalInternalReflection :18-32
      float cos_thetai = dot(neg(ray_dir), normal);
      float eta = ior;
      if (cos_thetai < 0.0f) eta = 1.0f / eta;
      float k2 = 1.0f - (1.0f - cos_thetai * cos_thetai) / (eta * eta);
      if (!(k2 < 0.0f)) {                                                     // refract :34-52
        f3 n = normal;
        f3 wo = neg(ray_dir);
        if (cos_thetai < 0.0f) { cos_thetai = -cos_thetai; n = neg(n); }
        float cos_theta2 = sqrtf(k2);
        nextDirection = normalize(neg(wo) * (1.0f / eta) - (cos_theta2 - cos_thetai / eta) * n);
      }
    }
    float cosThetaDiv = 1.0f / max2(fabsf(dot(nextDirection, normal)), 1.0e-20f);
    ms.color = bxdf * cosThetaDiv; ms.direction = nextDirection; ms.pdf = 1.0f; ms.pureSpecular = true;
  }
  return ms;
}

// EvalBxDF and EvalPDF for the light direction l (materials.adb:168-176, 217-226, 256-264, 333-341, 389-410);
// refl = reflect((-1.0)*v, n) with v = (-1.0)*ray_dir, i.e. reflect(ray_dir, n) -- the negations are exact
ART_D void mat_eval(const DMaterial& mat, f3 l, f3 n, f3 refl, f3& bxdf, float& pdf) {
  const int type = mat.type;
  if (type == MAT_LAMBERT) {
    bxdf = ld3(mat.color) * ART_INV_PI;
    pdf  = max2(dot(n, l), 0.0f) * ART_INV_PI;
  } else if (type == MAT_PHONG) {
    float cosPower = mat.param;
    float cosTheta = clampf(dot(l, refl), 0.0f, ART_PI * 0.499995f);
    float cosThetaDiv = 1.0f / max2(dot(l, n), 1.0e-20f);
    float pw = pow_ada_call(cosTheta, cosPower);
    bxdf = ld3(mat.color) * (cosPower + 2.0f) * 0.5f * ART_INV_PI * pw * cosThetaDiv;
    pdf  = pw * (cosPower + 1.0f) * (0.5f * ART_INV_PI);
  } else {
    bxdf = mk3(0.0f, 0.0f, 0.0f); pdf = 1.0f;
  }
}

}  // namespace art
