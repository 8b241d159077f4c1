// art_oracle.cpp -- CPU oracle: a restatement of the reference renderer's hot path.
//
// TEST INFRASTRUCTURE ONLY (see art_oracle.h).  PARITY UNPINNED: the reference has no
// tests/golden vectors and cannot be built here (no GNAT, no libembree3.a).  What holds it
// in place instead: tests/test_oracle_kat.py checks it bit for bit against two independent
// numpy binary32 restatements written from the Ada (tests/ada_numpy.py: rays and hits;
// tests/ada_numpy_shading.py: lights, BxDFs, sampling maps; tests/ada_numpy_integrators.py:
// DoPass and the three PathTrace bodies -- whole small frames, every accumulator word and
// the ray count), against the Random123 vectors and against closed-form radiometry.
//
// Every function cites the Ada source it follows (paths relative to the reference
// checkout).  Numerics contract: IEEE binary32 everywhere, no FMA
// contraction (build with -ffp-contract=off, SSE scalar math), left-to-right
// evaluation exactly as written in the Ada expressions, min/max as the explicit
// comparisons of generic_vector_math.adb:19-57, normalize = v * (1/sqrt(v.v)).
//
// Two things the reference takes from the GNAT run-time (not in the tree, version
// unknown) are replaced here, by design:
//   * Ada.Numerics.Float_Random (vector_math.adb:166-171) -> Philox4x32-10, counter
//     based: key = (seed, pixel index y*W+x), counter = (sample index, bounce, 0, 0);
//     the four 32-bit outputs are the bounce's uniforms in the order light u1, u2,
//     BSDF r1, r2 (u = (x >> 8) * 2^-24, in [0,1) like Float_Random).
//   * sin / cos / "**" (Generic_Elementary_Functions) -> deterministic binary32
//     kernels built from + - * / and sqrt only (det_sincos, det_log2, det_exp2 below,
//     <= ~2 ulp from the exact value), so that a GPU running the same arithmetic can be
//     compared bit for bit.  ao_use_libm(1) switches to libm's sinf/cosf/powf; the test
//     suite checks both variants agree at image level.
// Conditions on which the Ada code would raise Argument_Error (sqrt or ** of a negative
// value, which kills the worker task) are clamped to zero instead; each site is marked
// "ADA-RAISES".

#include "art_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#include <cfloat>

namespace {

// ---------------------------------------------------------------------------------------
// vector math -- generic_vector_math.adb / vector_math.adb
// ---------------------------------------------------------------------------------------

struct float3 { float x, y, z; };

const float INFINITY_ADA = FLT_MAX;                  // vector_math.ads:17  (Float'Last, not IEEE inf)
const float M_PI_F       = 3.14159265358979323846f;  // vector_math.ads:19  -> 0x40490FDB
const float INV_PI_F     = 0.318309886183790671538f; // vector_math.ads:20  -> 0x3EA2F983

inline float min2(float a, float b) { if (a < b) return a; else return b; }       // generic_vector_math.adb:19-26
inline float max2(float a, float b) { if (a >= b) return a; else return b; }      // :28-35
inline float max3(float a, float b, float c) {                                     // :48-57
  if (a >= b && a >= c) return a; else if (b >= c && b >= a) return b; else return c;
}
inline float clampf(float x, float a, float b) { return min2(max2(x, a), b); }    // :60-63

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // :81-88
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // :92-99
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // :101-108
inline float3 operator*(float3 a, float k)  { return {k * a.x, k * a.y, k * a.z}; }        // :148-151
inline float3 operator*(float k, float3 a)  { return {k * a.x, k * a.y, k * a.z}; }        // :153-156
inline float  dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        // :110-113
inline float3 cross(float3 a, float3 b) {                                                   // :115-122
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

inline float sign_ada(float x) { if (x >= 0.0f) return 1.0f; else return -1.0f; }  // vector_math.adb:49-56
inline float lerp_ada(float t, float a, float b) { return (1.0f - t) * a + t * b; } // :59-62

inline float3 normalize(float3 a) {                                                  // :64-72
  float l_inv = 1.0f / sqrtf(dot(a, a));
  return {l_inv * a.x, l_inv * a.y, l_inv * a.z};
}
inline float length(float3 a) { return sqrtf(dot(a, a)); }                           // :74-77
inline float3 reflect(float3 dir, float3 normal) {                                   // :79-82
  return normalize((normal * dot(dir, normal) * (-2.0f)) + dir);
}

struct float4x4 { float m[4][4]; };

inline float3 mul(const float4x4& M, float3 v) {                                     // :137-144
  float3 r;
  r.x = M.m[0][0] * v.x + M.m[0][1] * v.y + M.m[0][2] * v.z + M.m[0][3];
  r.y = M.m[1][0] * v.x + M.m[1][1] * v.y + M.m[1][2] * v.z + M.m[1][3];
  r.z = M.m[2][0] * v.x + M.m[2][1] * v.y + M.m[2][2] * v.z + M.m[2][3];
  return r;
}

float4x4 identity4() {
  float4x4 M; memset(&M, 0, sizeof(M));
  for (int i = 0; i < 4; i++) M.m[i][i] = 1.0f;
  return M;
}

float4x4 matmul(const float4x4& a, const float4x4& b) {                              // generic_vector_math.adb:233-256
  float4x4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
  return r;
}

float4x4 RotationMatrix(float angle, float3 a_v) {                                   // vector_math.adb:85-111
  float4x4 M = identity4();
  float3 v = normalize(a_v);
  float cos_t = cosf(angle);   // host-side, once per scene: libm (GNAT RTS in the reference)
  float sin_t = sinf(angle);
  M.m[0][0] = (1.0f - cos_t) * v.x * v.x + cos_t;
  M.m[0][1] = (1.0f - cos_t) * v.x * v.y - sin_t * v.z;
  M.m[0][2] = (1.0f - cos_t) * v.x * v.z + sin_t * v.y;
  M.m[1][0] = (1.0f - cos_t) * v.y * v.x + sin_t * v.z;
  M.m[1][1] = (1.0f - cos_t) * v.y * v.y + cos_t;
  M.m[1][2] = (1.0f - cos_t) * v.y * v.z - sin_t * v.x;
  M.m[2][0] = (1.0f - cos_t) * v.x * v.z - sin_t * v.y;
  M.m[2][1] = (1.0f - cos_t) * v.z * v.y + sin_t * v.x;
  M.m[2][2] = (1.0f - cos_t) * v.z * v.z + cos_t;
  return M;
}

// ---------------------------------------------------------------------------------------
// deterministic elementary functions (replace GNAT's Generic_Elementary_Functions)
// ---------------------------------------------------------------------------------------

std::atomic<int> g_use_libm{0};

inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// sin and cos of x (|x| < ~1e4): Cody-Waite reduction by pi/2 in three parts, then the
// classic single-precision minimax polynomials on [-pi/4, pi/4].
void det_sincos(float x, float* s_out, float* c_out) {
  float q  = x * 0.636619772367581343f;                      // x * 2/pi
  float qf = floorf(q + 0.5f);
  int   k  = (int)qf;
  float r  = x - qf * 1.5703125f;                            // pi/2 split: 8, 11 and 24 significant bits
  r = r - qf * 4.837512969970703125e-4f;
  r = r - qf * 7.54978995489188e-8f;
  float z  = r * r;
  float sp = -1.9515295891e-4f;
  sp = sp * z + 8.3321608736e-3f;
  sp = sp * z - 1.6666654611e-1f;
  float s  = sp * z * r + r;
  float cp = 2.443315711809948e-5f;
  cp = cp * z - 1.388731625493765e-3f;
  cp = cp * z + 4.166664568298827e-2f;
  float c  = cp * z * z - 0.5f * z + 1.0f;
  switch (k & 3) {
    case 0:  *s_out =  s; *c_out =  c; break;
    case 1:  *s_out =  c; *c_out = -s; break;
    case 2:  *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out =  s; break;
  }
}

// log2(x) for finite x > 0 (normal or denormal)
float det_log2(float x) {
  uint32_t ix = f2u(x);
  int e = 0;
  if (ix < 0x00800000u) { x = x * 8388608.0f; ix = f2u(x); e = -23; }   // denormal: scale by 2^23
  e += (int)(ix >> 23) - 127;
  uint32_t im = (ix & 0x007fffffu) | 0x3f800000u;                      // m in [1,2)
  float m = u2f(im);
  if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }                    // m in (sqrt(1/2), sqrt(2)]
  float s  = (m - 1.0f) / (m + 1.0f);
  float s2 = s * s;
  float p  = 0.111111111111f;                                          // 2*atanh(s) = 2s(1 + s^2/3 + s^4/5 + s^6/7 + s^8/9)
  p = p * s2 + 0.142857142857f;
  p = p * s2 + 0.2f;
  p = p * s2 + 0.333333333333f;
  p = p * s2 * s + s;
  float lnm = p + p;
  return (float)e + lnm * 1.44269504088896341f;
}

// 2^t
float det_exp2(float t) {
  if (!(t > -150.0f)) { if (t != t) return t; return 0.0f; }
  if (t > 128.0f) return INFINITY;
  float kf = floorf(t + 0.5f);
  int   k  = (int)kf;
  float z  = (t - kf) * 0.693147180559945309f;                          // |z| <= 0.3466
  float p  = 1.98412698413e-4f;                                        // 1/5040
  p = p * z + 1.38888888889e-3f;                                       // 1/720
  p = p * z + 8.33333333333e-3f;                                       // 1/120
  p = p * z + 4.16666666667e-2f;                                       // 1/24
  p = p * z + 1.66666666667e-1f;                                       // 1/6
  p = p * z + 0.5f;
  p = p * z + 1.0f;
  p = p * z + 1.0f;
  // scale by 2^k in two exact steps so that results down to the denormal range are produced by
  // ordinary IEEE multiplications
  int k1 = k / 2, k2 = k - k1;
  float s1 = u2f((uint32_t)(k1 + 127) << 23);
  float s2 = u2f((uint32_t)(k2 + 127) << 23);
  return p * s1 * s2;
}

// integer power by repeated squaring (the GNAT expansion of Float ** Integer; computed in
// binary32 here)
float det_powi(float x, int n) {
  float result = 1.0f, factor = x;
  int e = n;
  if (e != 0) {
    for (;;) {
      if (e & 1) result = result * factor;
      e = e / 2;
      if (e == 0) break;
      factor = factor * factor;
    }
  }
  return result;
}

// Left ** Right of Ada.Numerics.Generic_Elementary_Functions (GNAT a-ngelfu.adb, restated from
// its published algorithm: special cases, then for |Right| > 1 integer exponentiation of the
// integer part and square roots for the leading two fraction bits, the rest through exp/log).
float ada_expon(float left, float right) {
  if (g_use_libm.load(std::memory_order_relaxed)) {
    if (left <= 0.0f) return 0.0f;
    return powf(left, right);
  }
  if (left < 0.0f) return 0.0f;                 // ADA-RAISES Argument_Error
  if (left == 0.0f && right == 0.0f) return 1.0f; // ADA-RAISES Argument_Error
  if (right == 0.0f) return 1.0f;
  if (left == 0.0f) { if (right < 0.0f) return INFINITY; return 0.0f; } // ADA-RAISES Constraint_Error for right < 0
  if (left == 1.0f) return 1.0f;
  if (right == 1.0f) return left;
  if (right == 2.0f) return left * left;
  if (right == 0.5f) return sqrtf(left);
  float a_right = fabsf(right);
  if (a_right > 1.0f && a_right < 2147483648.0f) {
    int   int_part = (int)a_right;              // truncation
    float result   = det_powi(left, int_part);
    float rest     = a_right - (float)int_part;
    if (rest >= 0.5f) {
      float r1 = sqrtf(left);
      result = result * r1;
      rest   = rest - 0.5f;
      if (rest >= 0.25f) { result = result * sqrtf(r1); rest = rest - 0.25f; }
    } else if (rest >= 0.25f) {
      result = result * sqrtf(sqrtf(left));
      rest   = rest - 0.25f;
    }
    if (rest != 0.0f) result = result * det_exp2(rest * det_log2(left));
    if (right >= 0.0f) return result; else return 1.0f / result;
  }
  return det_exp2(right * det_log2(left));
}

// function pow, vector_math.adb:24-47
float pow_ada(float left, float right) {
  if (left == 0.0f && right == 0.0f) return 1.0f;   // ADA-RAISES Argument_Error
  else if (left < 0.0f) return 0.0f;                // ADA-RAISES Argument_Error
  else if (right == 0.0f) return 1.0f;
  else if (left == 0.0f) { if (right < 0.0f) return INFINITY; else return 0.0f; }
  else if (left == 1.0f) return 1.0f;
  else if (right == 1.0f) return left;
  else return ada_expon(left, right);
}

inline void sincos_ada(float x, float* s, float* c) {
  if (g_use_libm.load(std::memory_order_relaxed)) { *s = sinf(x); *c = cosf(x); }
  else det_sincos(x, s, c);
}

inline float sqrt_clamped(float x) { return sqrtf(max2(x, 0.0f)); }   // used only at ADA-RAISES sites

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants)
// ---------------------------------------------------------------------------------------

inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; r++) {
    if (r > 0) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    philox_round(c, k);
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

struct BounceRand { float u[4]; };   // light u1,u2 ; bsdf r1,r2  (vector_math.adb:166-171 rnd_uniform(0,1))

inline BounceRand bounce_rand(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce) {
  uint32_t ctr[4] = {sample, bounce, 0u, 0u}, key[2] = {seed, pixel}, o[4];
  philox4x32_10(ctr, key, o);
  BounceRand r;
  for (int i = 0; i < 4; i++) {
    float t = (float)(o[i] >> 8) * (1.0f / 16777216.0f);
    r.u[i] = 0.0f + (1.0f - 0.0f) * t;               // l + (h-l)*t
  }
  return r;
}

// ---------------------------------------------------------------------------------------
// scene -- scene.ads / scene.adb, geometry.ads
// ---------------------------------------------------------------------------------------

struct Ray { float3 origin{0, 0, 0}; float3 direction{0, 0, 1}; int x = 0, y = 0; };   // geometry.ads:15-19

struct Hit {                                                                            // geometry.ads:57-67
  int    prim_type = AO_PRIM_PLANE;
  bool   is_hit = false;
  float  t = INFINITY_ADA;
  float3 normal{0, 0, 0};
  int    mat = -1;        // material table index; -1 = null MaterialRef
  int    matId = 0;
  float  tx = 0, ty = 0;
  int    prim_index = -1;
};

Hit null_hit() {                                                                        // geometry.ads:126-135
  Hit h; h.prim_type = AO_PRIM_PLANE; h.prim_index = -1; h.is_hit = false; h.t = INFINITY_ADA;
  h.mat = -1; h.matId = 0; h.tx = 0; h.ty = 0; h.normal = {0, 1, 0};
  return h;
}

struct LiteGeomHit { bool is_hit = false; float tmin = 0, tmax = 0, u = 0, v = 0; };    // geometry.ads:106-111

struct AABB { float3 min{0, 0, 0}, max{0, 0, 0}; };

struct Mesh {                                                                           // geometry.ads:94-101
  std::vector<float3> vert_positions, vert_normals;
  std::vector<float>  tex_x, tex_y;
  std::vector<int>    tri;            // 3 per triangle
  std::vector<int>    material_ids;
  bool custom_material_ids = false;   // ao_set_mesh_tri_materials: per-triangle ids of a scene committed through the legacy seam
  AABB bbox;                          // zero-initialised: LoadMeshFromVSGF never resets it (geometry.adb:593-603)
};

} // namespace

struct ao_bvh;

struct ao_scene {                                                                        // scene.ads:61-73
  std::vector<ao_sphere>   spheres;
  ao_cornell               box;
  ao_flatlight             g_light;
  std::vector<ao_material> materials;
  ao_light                 light;
  ao_camera                cam;
  Mesh                     mymesh;
  ao_bvh*                  bvh = nullptr;   // lazily built for AO_MESH_NEAREST on big meshes
};

namespace {

inline float3 f3(const float* p) { return {p[0], p[1], p[2]}; }

// ---------------------------------------------------------------------------------------
// intersections -- geometry.adb
// ---------------------------------------------------------------------------------------

Hit IntersectAllSpheres(const Ray& r, const std::vector<ao_sphere>& sph) {               // geometry.adb:48-115
  float min_t = INFINITY_ADA;
  int   min_i = 0;
  float3 finalNormal{0, 1, 0};
  for (int i = 0; i < (int)sph.size(); i++) {
    float3 k = r.origin - f3(sph[i].pos);
    float  b = dot(k, r.direction);
    float  c = dot(k, k) - sph[i].r * sph[i].r;
    float  d = b * b - c;
    if (d >= 0.0f) {
      float sqrtd = sqrtf(d);
      float t1 = -b - sqrtd;
      float t2 = -b + sqrtd;
      if (t1 > 0.0f && t1 < min_t) { min_t = t1; min_i = i; }
      else if (t2 > 0.0f && t2 < min_t) { min_t = t2; min_i = i; }
    }
  }
  bool is_hit = (min_t > 0.0f && min_t < INFINITY_ADA);
  if (!is_hit) min_t = 1.0f;
  else finalNormal = normalize((r.origin + r.direction * min_t) - f3(sph[min_i].pos));
  Hit h;
  h.prim_type = AO_PRIM_SPHERE; h.prim_index = min_i; h.is_hit = is_hit; h.t = min_t;
  h.mat = sph.empty() ? -1 : sph[min_i].mat; h.matId = 0; h.normal = finalNormal; h.tx = 0; h.ty = 0;
  return h;
}

Hit IntersectFlatLight(const Ray& r, const ao_flatlight& lg, int lMat) {                  // geometry.adb:118-143
  float inv_dir_y = 1.0f / r.direction.y;
  float tmin = (lg.boxMax[1] - r.origin.y) * inv_dir_y;
  float3 hit_point = r.origin + tmin * r.direction;
  bool is_hit = (hit_point.x > lg.boxMin[0]) && (hit_point.x < lg.boxMax[0]) &&
                (hit_point.z > lg.boxMin[2]) && (hit_point.z < lg.boxMax[2]) && (tmin >= 0.0f);
  Hit h;
  h.prim_type = AO_PRIM_QUAD; h.prim_index = 0; h.is_hit = is_hit; h.t = tmin; h.mat = lMat; h.matId = 0;
  h.normal = {0, -1, 0}; h.tx = 0; h.ty = 0;
  return h;
}

LiteGeomHit IntersectBox(const Ray& r, const float3& bmin, const float3& bmax) {          // geometry.adb:146-191
  float inv_dir_x = 1.0f / r.direction.x;
  float inv_dir_y = 1.0f / r.direction.y;
  float inv_dir_z = 1.0f / r.direction.z;
  float lo  = (bmax.x - r.origin.x) * inv_dir_x;
  float hi  = (bmin.x - r.origin.x) * inv_dir_x;
  float lo1 = (bmax.y - r.origin.y) * inv_dir_y;
  float hi1 = (bmin.y - r.origin.y) * inv_dir_y;
  float lo2 = (bmax.z - r.origin.z) * inv_dir_z;
  float hi2 = (bmin.z - r.origin.z) * inv_dir_z;
  float tmin = min2(lo, hi);
  float tmax = max2(lo, hi);
  tmin = max2(tmin, min2(lo1, hi1));
  tmax = min2(tmax, max2(lo1, hi1));
  tmin = max2(tmin, min2(lo2, hi2));
  tmax = min2(tmax, max2(lo2, hi2));
  LiteGeomHit res;
  res.tmin = tmin; res.tmax = tmax;
  res.is_hit = (tmax > 0.0f) && (tmin <= tmax);
  return res;
}

Hit IntersectCornellBox(const Ray& r, const ao_cornell& bd) {                             // geometry.adb:193-229
  const float eps = 1.0e-5f;
  int planeId = 0;
  float3 bmin = f3(bd.bmin), bmax = f3(bd.bmax);
  LiteGeomHit tmpHit = IntersectBox(r, bmin, bmax);
  if (tmpHit.is_hit) {
    float3 p = r.origin + tmpHit.tmax * r.direction;
    if (fabsf(p.x - bmin.x) < eps) planeId = 0;
    if (fabsf(p.x - bmax.x) < eps) planeId = 1;
    if (fabsf(p.y - bmin.y) < eps) planeId = 2;
    if (fabsf(p.y - bmax.y) < eps) planeId = 3;
    if (fabsf(p.z - bmin.z) < eps) planeId = 4;
    if (fabsf(p.z - bmax.z) < eps) planeId = 5;
    Hit h;
    h.prim_type = AO_PRIM_PLANE; h.prim_index = planeId; h.is_hit = !(planeId == 5); h.t = tmpHit.tmax;
    h.mat = -1; h.matId = bd.mat_indices[planeId]; h.normal = f3(&bd.normals[3 * planeId]); h.tx = 0; h.ty = 0;
    return h;
  }
  return null_hit();
}

LiteGeomHit IntersectTriangle(const Ray& r, float3 A, float3 B, float3 C, float t_min, float t_max) { // geometry.adb:231-263
  const float epsilonDiv = 1.0e-25f;
  float3 edge1 = B - A;
  float3 edge2 = C - A;
  float3 pvec  = cross(r.direction, edge2);
  float3 tvec  = r.origin - A;
  float3 qvec  = cross(tvec, edge1);
  float invDet = 1.0f / max2(dot(edge1, pvec), epsilonDiv);
  float v = dot(tvec, pvec) * invDet;
  float u = dot(qvec, r.direction) * invDet;
  float t = dot(edge2, qvec) * invDet;
  LiteGeomHit res;
  res.is_hit = false;
  if (v > 0.0f && u > 0.0f && u + v < 1.0f && t > t_min && t < t_max) {
    res.u = u; res.v = v; res.tmin = t; res.tmax = t + 1.0e-6f; res.is_hit = true;
  }
  return res;
}

Hit mesh_hit_record(const Mesh& m, const LiteGeomHit& nearestHit, int nearestTriId) {      // geometry.adb:298-317
  int ia = m.tri[3 * nearestTriId], ib = m.tri[3 * nearestTriId + 1], ic = m.tri[3 * nearestTriId + 2];
  float w = 1.0f - nearestHit.u - nearestHit.v;
  float3 inorm = w * m.vert_normals[ia] + nearestHit.v * m.vert_normals[ib] + nearestHit.u * m.vert_normals[ic];
  float itx = w * m.tex_x[ia] + nearestHit.v * m.tex_x[ib] + nearestHit.u * m.tex_x[ic];
  float ity = w * m.tex_y[ia] + nearestHit.v * m.tex_y[ib] + nearestHit.u * m.tex_y[ic];
  Hit h;
  h.prim_type = AO_PRIM_TRIANGLE; h.prim_index = nearestTriId; h.is_hit = nearestHit.is_hit; h.t = nearestHit.tmin;
  h.mat = -1; h.matId = m.custom_material_ids ? m.material_ids[nearestTriId] : 2; h.normal = inorm; h.tx = itx; h.ty = ity;
  return h;
}

Hit IntersectMeshBF(const Ray& r, const Mesh& m) {                                         // geometry.adb:266-323
  if (m.tri.empty()) return null_hit();
  LiteGeomHit tmpHit = IntersectBox(r, m.bbox.min, m.bbox.max);
  if (tmpHit.is_hit) {
    LiteGeomHit nearestHit;
    nearestHit.tmin = 0.0f; nearestHit.tmax = 1000000.0f; nearestHit.is_hit = false;
    int nearestTriId = 0;
    int ntris = (int)m.tri.size() / 3;
    for (int i = 0; i < ntris; i++) {
      float3 A = m.vert_positions[m.tri[3 * i]];
      float3 B = m.vert_positions[m.tri[3 * i + 1]];
      float3 C = m.vert_positions[m.tri[3 * i + 2]];
      tmpHit = IntersectTriangle(r, A, B, C, nearestHit.tmin, nearestHit.tmax);
      if (tmpHit.is_hit) { nearestHit = tmpHit; nearestTriId = i; }   // NB: also moves the window to (t, t+1e-6)
    }
    return mesh_hit_record(m, nearestHit, nearestTriId);
  }
  return null_hit();
}

} // namespace

// --- CPU BVH for the nearest-hit mesh semantics -------------------------------
// Nearest front-facing hit with the reference's triangle test (window (tmin, tmax)), lowest
// triangle index on exact ties.  Median-split BVH over triangle centroids; not part of the
// reference (Embree is absent) -- "CPU BVH restatement, not Embree".

struct ao_bvh {
  struct Node { float bmin[3], bmax[3]; int left, right, first, count; };
  std::vector<Node> nodes;
  std::vector<int>  order;
  const float*      pos = nullptr;      // borrowed or owned copies below
  const int32_t*    tris = nullptr;
  std::vector<float>   pos_own;
  std::vector<int32_t> tris_own;
  int ntris = 0;
};

namespace {

void tri_bounds(const float* pos, const int32_t* tr, float mn[3], float mx[3]) {
  for (int a = 0; a < 3; a++) {
    float p0 = pos[3 * tr[0] + a], p1 = pos[3 * tr[1] + a], p2 = pos[3 * tr[2] + a];
    mn[a] = std::min(p0, std::min(p1, p2));
    mx[a] = std::max(p0, std::max(p1, p2));
  }
}

// Binned-SAH build (16 bins per axis, leaves of at most 4 triangles, median split where the bins cannot separate), single-ray
// traversal below: the "CPU BVH restatement, not Embree" of SURVEY 8d.  The tree only culls -- results are those of the triangle
// test and the tie rule -- so the split heuristic changes the speed of the CPU leg, never an answer.
inline float half_area(const float mn[3], const float mx[3]) {
  float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  return dx * dy + dy * dz + dz * dx;
}

void bvh_build_rec(ao_bvh& b, const std::vector<float>& tb, std::atomic<int>& next_node, int id, int first, int count, int depth) {
  // tb: per-triangle bounds, 6 floats (min.xyz, max.xyz); the centroid of a triangle is the middle of its box
  auto cen = [&](int t, int k) { return 0.5f * (tb[6 * (size_t)t + k] + tb[6 * (size_t)t + 3 + k]); };
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = first; i < first + count; i++) {
    const int t = b.order[i];
    for (int k = 0; k < 3; k++) {
      mn[k] = std::min(mn[k], tb[6 * (size_t)t + k]); mx[k] = std::max(mx[k], tb[6 * (size_t)t + 3 + k]);
      const float c = cen(t, k);
      cmn[k] = std::min(cmn[k], c); cmx[k] = std::max(cmx[k], c);
    }
  }
  for (int k = 0; k < 3; k++) { b.nodes[id].bmin[k] = mn[k]; b.nodes[id].bmax[k] = mx[k]; }
  b.nodes[id].first = first; b.nodes[id].count = count; b.nodes[id].left = b.nodes[id].right = -1;
  if (count <= 4) return;
  const int NB = 16;
  int cntb[3][NB]; float bmn[3][NB][3], bmx[3][NB][3]; float scale[3];
  for (int ax = 0; ax < 3; ax++) {
    const float ext = cmx[ax] - cmn[ax];
    scale[ax] = (ext > 0.0f) ? (float)NB / ext : 0.0f;
    for (int q = 0; q < NB; q++) { cntb[ax][q] = 0; for (int k = 0; k < 3; k++) { bmn[ax][q][k] = FLT_MAX; bmx[ax][q][k] = -FLT_MAX; } }
  }
  auto bin_of = [&](int t, int ax) { int q = (int)((cen(t, ax) - cmn[ax]) * scale[ax]); if (q >= NB) q = NB - 1; if (q < 0) q = 0; return q; };
  for (int i = first; i < first + count; i++) {
    const int t = b.order[i];
    for (int ax = 0; ax < 3; ax++) {
      if (scale[ax] == 0.0f) continue;
      const int q = bin_of(t, ax);
      cntb[ax][q]++;
      for (int k = 0; k < 3; k++) { bmn[ax][q][k] = std::min(bmn[ax][q][k], tb[6 * (size_t)t + k]); bmx[ax][q][k] = std::max(bmx[ax][q][k], tb[6 * (size_t)t + 3 + k]); }
    }
  }
  int best_axis = -1, best_split = 0; float best_cost = FLT_MAX;
  for (int ax = 0; ax < 3; ax++) {
    if (scale[ax] == 0.0f) continue;
    float la[NB], ra[NB]; int ln[NB], rn[NB];
    float amn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, amx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; int n = 0;
    for (int q = 0; q < NB; q++) { for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[ax][q][k]); amx[k] = std::max(amx[k], bmx[ax][q][k]); } n += cntb[ax][q]; ln[q] = n; la[q] = n ? half_area(amn, amx) : 0.0f; }
    for (int k = 0; k < 3; k++) { amn[k] = FLT_MAX; amx[k] = -FLT_MAX; } n = 0;
    for (int q = NB - 1; q >= 0; q--) { for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[ax][q][k]); amx[k] = std::max(amx[k], bmx[ax][q][k]); } n += cntb[ax][q]; rn[q] = n; ra[q] = n ? half_area(amn, amx) : 0.0f; }
    for (int q = 0; q + 1 < NB; q++) {
      if (ln[q] == 0 || rn[q + 1] == 0) continue;
      const float cost = la[q] * (float)ln[q] + ra[q + 1] * (float)rn[q + 1];
      if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = q; }
    }
  }
  int mid = first + count / 2;                               // no axis separates the centroids: split the list in half (keeps the leaves small)
  if (best_axis >= 0) {
    auto it = std::partition(b.order.begin() + first, b.order.begin() + first + count, [&](int t) { return bin_of(t, best_axis) <= best_split; });
    mid = (int)(it - b.order.begin());
    if (mid == first || mid == first + count) mid = first + count / 2;
  }
  b.nodes[id].count = 0;
  const int l = next_node.fetch_add(2), r = l + 1;
  b.nodes[id].left = l; b.nodes[id].right = r;
  if (depth < 5 && count > 50000) {                          // the top of a big tree: the two halves are independent
    std::thread th([&]() { bvh_build_rec(b, tb, next_node, l, first, mid - first, depth + 1); });
    bvh_build_rec(b, tb, next_node, r, mid, first + count - mid, depth + 1);
    th.join();
  } else {
    bvh_build_rec(b, tb, next_node, l, first, mid - first, depth + 1);
    bvh_build_rec(b, tb, next_node, r, mid, first + count - mid, depth + 1);
  }
}

// conservative slab test (double precision, padded) -- only a culling device; exactness of the
// result comes from the triangle test and the tie rule.
inline bool slab(const ao_bvh::Node& n, const double o[3], const double inv[3], double tmax_cur) {
  double t0 = 0.0, t1 = tmax_cur;
  for (int a = 0; a < 3; a++) {
    double lo = ((double)n.bmin[a] - o[a]) * inv[a];
    double hi = ((double)n.bmax[a] - o[a]) * inv[a];
    if (lo > hi) std::swap(lo, hi);
    if (lo != lo || hi != hi) continue;          // 0*inf: axis-parallel ray on a slab plane -> do not cull
    if (std::isfinite(lo)) lo -= 1e-4 * (fabs(lo) + 1.0);   // inf - inf would be NaN: an axis-parallel ray BESIDE the slab must still be culled
    if (std::isfinite(hi)) hi += 1e-4 * (fabs(hi) + 1.0);
    if (lo > t0) t0 = lo;
    if (hi < t1) t1 = hi;
  }
  return t0 <= t1;
}

void bvh_closest_one(const ao_bvh& b, const Ray& r, float tmin, float tmax, int* prim, float* t_out, float* u_out, float* v_out) {
  int best = -1; float best_t = tmax, bu = 0, bv = 0;
  if (!b.nodes.empty()) {
    double o[3] = {r.origin.x, r.origin.y, r.origin.z};
    double inv[3] = {1.0 / (double)r.direction.x, 1.0 / (double)r.direction.y, 1.0 / (double)r.direction.z};
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
      const ao_bvh::Node& n = b.nodes[stack[--sp]];
      if (!slab(n, o, inv, (double)best_t * (1.0 + 1e-6) + 1e-6)) continue;
      if (n.left < 0) {
        for (int i = n.first; i < n.first + n.count; i++) {
          int t = b.order[i];
          const int32_t* tr = b.tris + 3 * t;
          // window (tmin, +inf): ties with the current best are resolved below by index
          LiteGeomHit h = IntersectTriangle(r, f3(b.pos + 3 * tr[0]), f3(b.pos + 3 * tr[1]), f3(b.pos + 3 * tr[2]), tmin, tmax);
          if (h.is_hit && (h.tmin < best_t || (h.tmin == best_t && best >= 0 && t < best))) { best = t; best_t = h.tmin; bu = h.u; bv = h.v; }
        }
      } else if (sp < 126) {
        // nearer child (along the ray's dominant axis) on top of the stack
        const ao_bvh::Node& L = b.nodes[n.left]; const ao_bvh::Node& R = b.nodes[n.right];
        int ax = 0; double am = fabs(inv[0]) < 1e300 ? fabs(1.0 / inv[0]) : 0.0;
        for (int k = 1; k < 3; k++) { double v = fabs(inv[k]) < 1e300 ? fabs(1.0 / inv[k]) : 0.0; if (v > am) { am = v; ax = k; } }
        const bool left_first = (inv[ax] >= 0.0) ? (L.bmin[ax] <= R.bmin[ax]) : (L.bmax[ax] >= R.bmax[ax]);
        if (left_first) { stack[sp++] = n.right; stack[sp++] = n.left; } else { stack[sp++] = n.left; stack[sp++] = n.right; }
      }
    }
  }
  *prim = best; if (t_out) *t_out = best_t; if (u_out) *u_out = bu; if (v_out) *v_out = bv;
}

void bruteforce_closest_one(const float* pos, const int32_t* tris, int ntris, const Ray& r, float tmin, float tmax,
                            int* prim, float* t_out, float* u_out, float* v_out) {
  int best = -1; float best_t = tmax, bu = 0, bv = 0;
  for (int t = 0; t < ntris; t++) {
    const int32_t* tr = tris + 3 * t;
    LiteGeomHit h = IntersectTriangle(r, f3(pos + 3 * tr[0]), f3(pos + 3 * tr[1]), f3(pos + 3 * tr[2]), tmin, tmax);
    if (h.is_hit && h.tmin < best_t) { best = t; best_t = h.tmin; bu = h.u; bv = h.v; }   // index order => lowest index wins ties
  }
  *prim = best; if (t_out) *t_out = best_t; if (u_out) *u_out = bu; if (v_out) *v_out = bv;
}

Hit IntersectMeshNearest(const Ray& r, const ao_scene& scn) {
  const Mesh& m = scn.mymesh;
  if (m.tri.empty()) return null_hit();
  int prim; float t, u, v;
  int ntris = (int)m.tri.size() / 3;
  if (scn.bvh) bvh_closest_one(*scn.bvh, r, 0.0f, 1000000.0f, &prim, &t, &u, &v);
  else bruteforce_closest_one(&m.vert_positions[0].x, m.tri.data(), ntris, r, 0.0f, 1000000.0f, &prim, &t, &u, &v);
  if (prim < 0) return null_hit();
  LiteGeomHit lh; lh.is_hit = true; lh.tmin = t; lh.tmax = t + 1.0e-6f; lh.u = u; lh.v = v;
  return mesh_hit_record(m, lh, prim);
}

// ---------------------------------------------------------------------------------------
// Scene.Find_Closest_Hit -- scene.adb:56-86
// ---------------------------------------------------------------------------------------

struct Ctx { const ao_scene* scn; int mesh_mode; uint64_t rays = 0; };

Hit Find_Closest_Hit(Ctx& cx, const Ray& r) {
  const ao_scene& s = *cx.scn;
  cx.rays++;
  Hit hits[4];
  int nearestHitIndex = 0;
  float nearestHitDist = INFINITY_ADA;
  hits[0] = IntersectAllSpheres(r, s.spheres);
  hits[1] = IntersectCornellBox(r, s.box);
  if (s.light.shape == AO_LIGHT_RECT) hits[2] = IntersectFlatLight(r, s.g_light, 4);
  hits[3] = (cx.mesh_mode == AO_MESH_NEAREST) ? IntersectMeshNearest(r, s) : IntersectMeshBF(r, s.mymesh);
  for (int i = 0; i < 4; i++) {
    if (hits[i].is_hit && hits[i].t < nearestHitDist) { nearestHitIndex = i; nearestHitDist = hits[i].t; }
  }
  if (hits[nearestHitIndex].mat < 0) hits[nearestHitIndex].mat = hits[nearestHitIndex].matId;
  return hits[nearestHitIndex];
}

// ---------------------------------------------------------------------------------------
// Lights -- lights.adb
// ---------------------------------------------------------------------------------------

struct ShadowSample { float3 pos{0, 0, 0}, dir{0, 0, 0}, intensity{0, 0, 0}; float pdf = 1.0f; };   // lights.ads:15-20

const float epsilonDiv = 1.0e-20f;                                                                   // lights.adb:39

float PdfAtoW(float aPdfA, float aDist, float aCosThere) { return aPdfA * aDist * aDist / max2(aCosThere, epsilonDiv); } // :42-45

void CoordinateSystem(float3 v1, float3* v2, float3* v3) {                                            // :107-121
  if (fabsf(v1.x) > fabsf(v1.y)) {
    float invLen = 1.0f / sqrtf(v1.x * v1.x + v1.z * v1.z);
    *v2 = {-v1.z * invLen, 0.0f, v1.x * invLen};
  } else {
    float invLen = 1.0f / sqrtf(v1.y * v1.y + v1.z * v1.z);
    *v2 = {0.0f, v1.z * invLen, -v1.y * invLen};
  }
  *v3 = cross(v1, *v2);
}

float DistanceSquared(float3 a, float3 b) { float3 diff = b - a; return dot(diff, diff); }           // :123-130

float3 UniformSampleSphere(float u1, float u2) {                                                      // :133-142
  float z = 1.0f - 2.0f * u1;
  float r = sqrtf(max2(0.0f, 1.0f - z * z));
  float phi = 2.0f * M_PI_F * u2;
  float s, c; sincos_ada(phi, &s, &c);
  return {r * c, r * s, z};
}

float3 UniformSampleCone(float u1, float u2, float costhetamax, float3 x, float3 y, float3 z) {      // :144-151
  float costheta = lerp_ada(u1, costhetamax, 1.0f);
  float sintheta = sqrt_clamped(1.0f - costheta * costheta);                                          // ADA-RAISES if costheta rounds above 1
  float phi = u2 * 2.0f * M_PI_F;
  float s, c; sincos_ada(phi, &s, &c);
  return c * sintheta * x + s * sintheta * y + costheta * z;
}

float UniformConePdf(float cosThetaMax) { return 1.0f / (2.0f * M_PI_F * (1.0f - cosThetaMax)); }    // :153-159

void RaySphereIntersect(float3 rayPos, float3 rayDir, float3 sphPos, float radius, float* rx, float* ry) { // :161-194
  float3 k = rayPos - sphPos;
  float b = dot(k, rayDir);
  float c = dot(k, k) - radius * radius;
  float d = b * b - c;
  if (d >= 0.0f) {
    float sqrtd = sqrtf(d);
    float t1 = -b - sqrtd, t2 = -b + sqrtd;
    *rx = min2(t1, t2); *ry = max2(t1, t2);
  } else { *rx = -INFINITY_ADA; *ry = -INFINITY_ADA; }
}

float Light_EvalPDF(const ao_light& l, float3 lluminatingPoint, float3 rayDir, float hitDist) {
  if (l.shape == AO_LIGHT_RECT) {                                                                     // :81-85
    float cosTheta = max2(dot(rayDir, (-1.0f) * f3(l.normal)), 0.0f);
    return PdfAtoW(1.0f / l.surfaceArea, hitDist, cosTheta);
  }
  // SphereLight :242-255
  float3 center = f3(l.center);
  if (DistanceSquared(lluminatingPoint, center) - l.radius * l.radius < 1.0e-4f) return 1.0f / l.surfaceArea;
  float sinThetaMax2 = l.radius * l.radius / DistanceSquared(lluminatingPoint, center);
  float cosThetaMax  = sqrtf(max2(0.0f, 1.0f - sinThetaMax2));
  return UniformConePdf(cosThetaMax);
}

ShadowSample Light_Sample(const ao_light& l, const BounceRand& gen, float3 lluminatingPoint) {
  ShadowSample res;
  if (l.shape == AO_LIGHT_RECT) {                                                                     // :56-79
    float r1 = gen.u[0], r2 = gen.u[1];
    res.pos.x = l.boxMin[0] + r1 * (l.boxMax[0] - l.boxMin[0]);
    res.pos.y = l.boxMin[1];
    res.pos.z = l.boxMin[2] + r2 * (l.boxMax[2] - l.boxMin[2]);
    res.dir = f3(l.normal);
    float3 rayDir = res.pos - lluminatingPoint;
    float d = length(rayDir);
    rayDir = rayDir * (1.0f / d);
    float cosTheta = max2(dot(rayDir, (-1.0f) * f3(l.normal)), 0.0f);
    res.pdf = PdfAtoW(1.0f / l.surfaceArea, d, cosTheta);
    res.intensity = f3(l.intensity);
    return res;
  }
  // SphereLight :197-240
  float u1 = gen.u[0], u2 = gen.u[1];
  float3 center = f3(l.center);
  res.intensity = f3(l.intensity);
  if (DistanceSquared(lluminatingPoint, center) - l.radius * l.radius < 1.0e-4f) {
    res.pos = center + l.radius * UniformSampleSphere(u1, u2);
    res.dir = normalize(res.pos - center);
    return res;                                    // pdf stays at its default 1.0 (lights.ads:19)
  }
  float3 wc = normalize(center - lluminatingPoint), wcX, wcY;
  CoordinateSystem(wc, &wcX, &wcY);
  float sinThetaMax2 = l.radius * l.radius / DistanceSquared(lluminatingPoint, center);
  float cosThetaMax  = sqrtf(max2(0.0f, 1.0f - sinThetaMax2));
  float3 rdir = UniformSampleCone(u1, u2, cosThetaMax, wcX, wcY, wc);
  float3 rpos = lluminatingPoint + rdir * (1.0e-3f);
  float hx, hy;
  RaySphereIntersect(rpos, rdir, center, l.radius, &hx, &hy);
  float thit;
  if (hx < 0.0f) thit = dot(center - lluminatingPoint, normalize(rdir));
  else thit = hx;
  res.pos = rpos + thit * rdir;
  res.dir = normalize(res.pos - center);
  res.pdf = Light_EvalPDF(l, lluminatingPoint, rdir, thit);
  return res;
}

// ---------------------------------------------------------------------------------------
// sampling maps -- vector_math.adb:175-326
// ---------------------------------------------------------------------------------------

float3 GetPerpendicular(float3 a_vec) {                                                               // :175-200
  float3 leastPerpendicular;
  float xProjection = fabsf(a_vec.x), yProjection = fabsf(a_vec.y), zProjection = fabsf(a_vec.z);
  if ((xProjection <= yProjection + 1.0e-5f) && (xProjection <= zProjection + 1.0e-5f)) leastPerpendicular = {1, 0, 0};
  else if ((yProjection < xProjection + 1.0e-5f) && (yProjection <= zProjection + 1.0e-5f)) leastPerpendicular = {0, 1, 0};
  else leastPerpendicular = {0, 0, 1};
  return normalize(cross(a_vec, leastPerpendicular));
}

float3 MapTail(float3 deviation, float3 direction, float3 normal) {                                   // :218-250 == :278-310
  float3 ny = direction;
  float3 nx = GetPerpendicular(ny);
  float3 nz = normalize(cross(nx, ny));
  float3 tmp = ny; ny = nz; nz = tmp;
  float3 res = nx * deviation.x + ny * deviation.y + nz * deviation.z;
  float invSign;
  if (dot(direction, normal) >= 0.0f) invSign = 1.0f; else invSign = -1.0f;
  if (invSign * dot(res, normal) < 0.0f) {
    nx = normalize(cross(normal, direction));
    nz = normalize(cross(nx, ny));
    if (dot(nz, res) < 0.0f) nz = (-1.0f) * nz;
    res = reflect((-1.0f) * res, nz);
    if (dot(res, normal) < 0.0f) res = direction;
  }
  return res;
}

float3 MapSampleToCosineDist(float r1, float r2, float3 direction, float3 normal, float power) {     // :202-252
  float e = power;
  float sin_phi, cos_phi;
  sincos_ada(2.0f * r1 * M_PI_F, &sin_phi, &cos_phi);
  float cos_theta = ada_expon(1.0f - r2, 1.0f / (e + 1.0f));
  float sin_theta = sqrt_clamped(1.0f - cos_theta * cos_theta);                                       // ADA-RAISES if negative
  float3 deviation{sin_theta * cos_phi, sin_theta * sin_phi, cos_theta};
  return MapTail(deviation, direction, normal);
}

float3 MapSampleToCosineDistFixed(float r1, float r2, float3 direction, float3 normal, float power) { // :255-312
  float h = sqrt_clamped(1.0f - ada_expon(r1, 2.0f / (power + 1.0f)));                                // ADA-RAISES if negative
  float s, c;
  sincos_ada(2.0f * M_PI_F * r2, &s, &c);
  float3 deviation;
  deviation.x = h * c;
  deviation.y = h * s;
  deviation.z = ada_expon(r1, 1.0f / (power + 1.0f));
  return MapTail(deviation, direction, normal);
}

// ---------------------------------------------------------------------------------------
// Materials -- materials.adb
// ---------------------------------------------------------------------------------------

struct MatSample { float3 color{0, 0, 0}, direction{0, 0, 0}; float pdf = 1.0f; bool pureSpecular = false; }; // materials.ads:18-23

const float Eps_Div = 1.0e-20f, Eps_Cos = 1.0e-6f;                                                    // materials.adb:15-16

bool TotalInternalReflection(float ior, float3 rayDir, float3 normal) {                               // :18-32
  float cos_thetai = dot((-1.0f) * rayDir, normal);
  float eta = ior;
  if (cos_thetai < 0.0f) eta = 1.0f / eta;
  return (1.0f - (1.0f - cos_thetai * cos_thetai) / (eta * eta)) < 0.0f;
}

float3 refract_ada(float ior, float3 rayDir, float3 normal) {                                         // :34-52
  float3 n = normal;
  float3 wo = (-1.0f) * rayDir;
  float cos_thetai = dot((-1.0f) * rayDir, normal);
  float eta = ior;
  if (cos_thetai < 0.0f) { eta = 1.0f / eta; cos_thetai = -cos_thetai; n = (-1.0f) * n; }
  float cos_theta2 = sqrtf(1.0f - (1.0f - cos_thetai * cos_thetai) / (eta * eta));   // caller guarantees >= 0 (no TIR)
  return normalize((-1.0f) * wo * (1.0f / eta) - (cos_theta2 - cos_thetai / eta) * n);
}

float fresnelDielectric(float cosTheta1, float cosTheta2, float etaExt, float etaInt) {               // :70-76
  float Rs = (etaExt * cosTheta1 - etaInt * cosTheta2) / (etaExt * cosTheta1 + etaInt * cosTheta2);
  float Rp = (etaInt * cosTheta1 - etaExt * cosTheta2) / (etaInt * cosTheta1 + etaExt * cosTheta2);
  return (Rs * Rs + Rp * Rp) / 2.0f;
}

float fresnel(float cosTheta1, float a_etaExt, float a_etaInt) {                                      // :79-99
  float etaExt = a_etaExt, etaInt = a_etaInt;
  if (cosTheta1 < 0.0f) { float tmp = etaExt; etaExt = etaInt; etaInt = tmp; }
  float sinTheta2 = (etaExt / etaInt) * sqrtf(max2(0.0f, 1.0f - cosTheta1 * cosTheta1));
  if (sinTheta2 > 1.0f) return 1.0f;
  float cosTheta2 = sqrtf(max2(0.0f, 1.0f - sinTheta2 * sinTheta2));
  return fresnelDielectric(fabsf(cosTheta1), cosTheta2, etaInt, etaExt);
}

bool Mat_IsLight(const ao_material& m) { return m.type == AO_MAT_LIGHT; }

float3 Mat_Emittance(const ao_scene& s, const ao_material& m) {                                       // :147-156
  if (m.type == AO_MAT_LIGHT) { if (m.light < 0) return {0, 0, 0}; return f3(s.light.intensity); }
  return {0, 0, 0};
}

MatSample Mat_SampleAndEvalBxDF(const ao_material& mat, const BounceRand& gen, float3 ray_dir, float3 normal) {
  MatSample ms;
  switch (mat.type) {
    case AO_MAT_LAMBERT: {                                                                            // :197-215
      float3 newDir = MapSampleToCosineDist(gen.u[2], gen.u[3], normal, normal, 1.0f);                // vector_math.adb:314-319
      float cosTheta = dot(newDir, normal);
      float pdf = fabsf(cosTheta) * INV_PI_F;
      float3 color = f3(mat.color) * INV_PI_F;
      if (cosTheta < Eps_Cos) color = {0, 0, 0};
      ms.color = color; ms.direction = newDir; ms.pdf = pdf; ms.pureSpecular = false;
      return ms;
    }
    case AO_MAT_MIRROR: {                                                                             // :247-254
      float3 nextDir = reflect(ray_dir, normal);
      float cosThetaDiv = 1.0f / max2(dot(nextDir, normal), Eps_Div);
      ms.color = f3(mat.color) * cosThetaDiv; ms.direction = nextDir; ms.pdf = 1.0f; ms.pureSpecular = true;
      return ms;
    }
    case AO_MAT_GLASS: {                                                                              // :285-331
      float3 refl = f3(mat.color), trans = f3(mat.color2);
      float f = fresnel(dot(ray_dir, normal), mat.param, 1.0f);                                       // ApplyFresnel :285-293
      refl = f * refl;
      trans = (1.0f - f) * trans;
      float ksitrans = length(trans) / (length(refl) + length(trans));
      float ksirefl  = length(refl) / (length(refl) + length(trans));
      float ksi = gen.u[2];
      float3 nextDirection, bxdf;
      if (ksi > ksitrans) {
        nextDirection = reflect(ray_dir, normal);
        bxdf = refl * (1.0f / ksirefl);
      } else {
        bxdf = trans * (1.0f / ksitrans);
        if (!TotalInternalReflection(mat.param, ray_dir, normal)) nextDirection = refract_ada(mat.param, ray_dir, normal);
        else nextDirection = reflect(ray_dir, normal);
      }
      float cosThetaDiv = 1.0f / max2(fabsf(dot(nextDirection, normal)), Eps_Div);
      ms.color = bxdf * cosThetaDiv; ms.direction = nextDirection; ms.pdf = 1.0f; ms.pureSpecular = true;
      return ms;
    }
    case AO_MAT_PHONG: {                                                                              // :363-387
      float3 r = reflect(ray_dir, normal);
      float3 nextDir = MapSampleToCosineDistFixed(gen.u[2], gen.u[3], r, normal, mat.param);          // vector_math.adb:321-326
      float cosTheta = clampf(dot(nextDir, r), 0.0f, M_PI_F * 0.499995f);
      float3 color = f3(mat.color) * (mat.param + 2.0f) * 0.5f * INV_PI_F * pow_ada(cosTheta, mat.param);
      float pdf = pow_ada(cosTheta, mat.param) * (mat.param + 1.0f) * (0.5f * INV_PI_F);
      float cosThetaGeo = dot(nextDir, normal);
      float cosThetaDiv = 1.0f / max2(fabsf(cosThetaGeo), Eps_Div);
      if (cosThetaGeo < Eps_Cos) color = {0, 0, 0};
      ms.color = color * cosThetaDiv; ms.direction = nextDir; ms.pdf = pdf; ms.pureSpecular = false;
      return ms;
    }
    default:                                                                                          // MaterialLight :163-166 (and null)
      ms.color = {0, 0, 0}; ms.direction = {0, 0, 0}; ms.pdf = 1.0f; ms.pureSpecular = false;
      return ms;
  }
}

float3 Mat_EvalBxDF(const ao_material& mat, float3 l, float3 v, float3 n) {
  switch (mat.type) {
    case AO_MAT_LAMBERT: return f3(mat.color) * INV_PI_F;                                             // :217-220
    case AO_MAT_PHONG: {                                                                              // :389-398
      float3 r = reflect((-1.0f) * v, n);
      float cosTheta = clampf(dot(l, r), 0.0f, M_PI_F * 0.499995f);
      float cosThetaDiv = 1.0f / max2(dot(l, n), Eps_Div);
      return f3(mat.color) * (mat.param + 2.0f) * 0.5f * INV_PI_F * pow_ada(cosTheta, mat.param) * cosThetaDiv;
    }
    default: return {0, 0, 0};                                                                        // light :168-171, mirror :256-259, glass :333-336
  }
}

float Mat_EvalPDF(const ao_material& mat, float3 l, float3 v, float3 n) {
  switch (mat.type) {
    case AO_MAT_LAMBERT: { float cosTheta = max2(dot(n, l), 0.0f); return cosTheta * INV_PI_F; }      // :222-226
    case AO_MAT_PHONG: {                                                                              // :401-410
      float3 r = reflect((-1.0f) * v, n);
      float cosTheta = clampf(dot(l, r), 0.0f, M_PI_F * 0.499995f);
      return pow_ada(cosTheta, mat.param) * (mat.param + 1.0f) * (0.5f * INV_PI_F);
    }
    default: return 1.0f;                                                                             // :173-176, :261-264, :338-341
  }
}

// ---------------------------------------------------------------------------------------
// Ray_Tracer -- ray_tracer.adb, ray_tracer-integrators.adb
// ---------------------------------------------------------------------------------------

const float G_Epsilon = 1.0e-5f, G_Epsilon_Div = 1.0e-20f;                                            // ray_tracer.ads:29-30

bool Compute_Shadow(Ctx& cx, float3 hit_pos, float3 lpos) {                                           // ray_tracer.adb:100-132
  float epsilon = max3(fabsf(hit_pos.x), fabsf(hit_pos.y), fabsf(hit_pos.z)) * 0.000000001f;
  epsilon = max2(epsilon, 1.0e-30f);
  Ray shadowRay;
  shadowRay.direction = normalize(lpos - hit_pos);
  shadowRay.origin    = hit_pos + shadowRay.direction * epsilon;
  Hit h = Find_Closest_Hit(cx, shadowRay);
  float maxDist  = length(hit_pos - lpos);
  float epsilon2 = max2(maxDist * 0.000001f, 1.0e-30f);
  return h.is_hit && (h.t < maxDist - epsilon2 && h.t > 10.0f * epsilon);
}

float3 EyeRayDirection(int width, int height, int x, int y, float off_x, float off_y) {               // ray_tracer.adb:61-97
  const float fov = 1.57079632679489661923f;                 // Pi / 2.0
  float3 res;
  res.x = (float)x + off_x - ((float)width / 2.0f);
  res.y = (float)y + off_y - ((float)height / 2.0f);
  res.z = -(float)width / tanf(fov / 2.0f);                  // safe_tan (vector_math.adb:14-22); tanf(0.78539819f) == 1.0f
  return normalize(res);
}

const float AA_OFF[4][2] = {{1.0f / 3.0f, 1.0f / 3.0f}, {1.0f / 3.0f, 2.0f / 3.0f},                  // ray_tracer.adb:76-91
                            {2.0f / 3.0f, 1.0f / 3.0f}, {2.0f / 3.0f, 2.0f / 3.0f}};

Ray make_eye_ray(const ao_scene& s, int width, int height, int x, int y, int sub) {                   // integrators.adb:37-46, 56-58
  Ray r; r.x = x; r.y = y;
  r.origin = f3(s.cam.pos);
  float4x4 M; memcpy(&M, s.cam.matrix, sizeof(M));
  float3 d = (sub < 0) ? EyeRayDirection(width, height, x, y, 0.5f, 0.5f)
                       : EyeRayDirection(width, height, x, y, AA_OFF[sub][0], AA_OFF[sub][1]);
  r.direction = normalize(mul(M, d));
  return r;
}

struct PathCtx { Ctx cx; uint32_t seed, pixel, sample; int max_depth; int render_type; };

const MatSample StartSample = {{0, 0, 0}, {0, 0, 0}, 1.0f, true};                                     // materials.ads:25

float3 PathTrace(PathCtx& pc, const Ray& r, const MatSample& prevSample, int recursion_level) {
  const ao_scene& scn = *pc.cx.scn;
  float3 explicitColor{0, 0, 0};
  if (recursion_level == 0) return {0, 0, 0};                                                         // integrators.adb:92,145,212
  Hit h = Find_Closest_Hit(pc.cx, r);
  if (!h.is_hit) return {0, 0, 0};                                                                    // :98,151,218
  const ao_material& hmat = scn.materials[h.mat];
  BounceRand gen = bounce_rand(pc.seed, pc.pixel, pc.sample, (uint32_t)(pc.max_depth - recursion_level));

  if (Mat_IsLight(hmat)) {
    if (pc.render_type == AO_PT_SHADOW) return {0, 0, 0};                                             // :155-157
    if (dot((-1.0f) * r.direction, h.normal) < 0.0f) return {0, 0, 0};                                // :103,224
    if (pc.render_type == AO_PT_STUPID) return Mat_Emittance(scn, hmat);                              // :106
    float lgtPdf  = Light_EvalPDF(scn.light, r.origin, r.direction, h.t);                             // :231
    float bsdfPdf = prevSample.pdf;
    float misWeight;
    if (prevSample.pureSpecular) misWeight = 1.0f;
    else misWeight = bsdfPdf * bsdfPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf);
    return Mat_Emittance(scn, hmat) * misWeight;                                                      // :242
  }

  if (pc.render_type == AO_PT_SHADOW) {                                                               // :161-178
    float3 hpos = (r.origin + r.direction * h.t);
    ShadowSample lsam = Light_Sample(scn.light, gen, hpos);
    float3 sdir = normalize(lsam.pos - hpos);
    if (!Compute_Shadow(pc.cx, hpos, lsam.pos)) {
      float3 bxdfVal = Mat_EvalBxDF(hmat, sdir, (-1.0f) * r.direction, h.normal);
      float cosTheta1 = max2(dot(sdir, h.normal), 0.0f);
      explicitColor = lsam.intensity * (cosTheta1 * bxdfVal) * (1.0f / max2(lsam.pdf, G_Epsilon_Div));
    }
  } else if (pc.render_type == AO_PT_MIS) {                                                           // :251-287
    float3 hpos = (r.origin + r.direction * h.t);
    ShadowSample lsam = Light_Sample(scn.light, gen, hpos);
    float3 sdir = normalize(lsam.pos - hpos);
    float lgtPdf = lsam.pdf;
    if (!Compute_Shadow(pc.cx, hpos, lsam.pos)) {
      float bsdfPdf = Mat_EvalPDF(hmat, sdir, (-1.0f) * r.direction, h.normal);
      float3 bxdfVal = Mat_EvalBxDF(hmat, sdir, (-1.0f) * r.direction, h.normal);
      float cosTheta1 = max2(dot(sdir, h.normal), 0.0f);
      float misWeight = lgtPdf * lgtPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf);
      explicitColor = lsam.intensity * (1.0f / max2(lgtPdf, G_Epsilon_Div)) * (cosTheta1 * bxdfVal) * misWeight;
    } else explicitColor = {0, 0, 0};
  }

  MatSample matSam = Mat_SampleAndEvalBxDF(hmat, gen, r.direction, h.normal);                         // :116,183,291
  float3 bxdfVal = matSam.color * (1.0f / max2(matSam.pdf, G_Epsilon_Div));
  float cosTheta = dot(matSam.direction, h.normal);
  Ray nextRay = r;
  nextRay.origin    = r.origin + r.direction * h.t;
  nextRay.direction = matSam.direction;
  nextRay.origin    = nextRay.origin + sign_ada(cosTheta) * h.normal * G_Epsilon;
  float3 next = PathTrace(pc, nextRay, matSam, recursion_level - 1);
  if (pc.render_type == AO_PT_STUPID) return fabsf(cosTheta) * bxdfVal * next;                        // :124
  return explicitColor + (fabsf(cosTheta) * bxdfVal) * next;                                          // :191,299
}

const float3 PALETTE[8] = {{0.5f, 0, 0}, {0, 0.5f, 0}, {0, 0, 0.5f}, {0.5f, 0.5f, 0.5f},              // ray_tracer.adb:210-211
                           {0.5f, 0.5f, 0}, {0.5f, 0, 0.5f}, {0, 0.5f, 0.5f}, {0.75f, 0.75f, 0.75f}};

void hit_to_c(const Hit& h, ao_hit* o) {
  o->is_hit = h.is_hit ? 1 : 0; o->prim_type = h.prim_type; o->prim_index = h.prim_index; o->matId = h.matId; o->mat = h.mat;
  o->t = h.t; o->normal[0] = h.normal.x; o->normal[1] = h.normal.y; o->normal[2] = h.normal.z; o->tx = h.tx; o->ty = h.ty;
}

void ensure_bvh(const ao_scene* cs, int mesh_mode) {
  ao_scene* s = const_cast<ao_scene*>(cs);
  if (mesh_mode == AO_MESH_NEAREST && !s->bvh && s->mymesh.tri.size() / 3 > 64)
    s->bvh = ao_bvh_build(&s->mymesh.vert_positions[0].x, (int)s->mymesh.vert_positions.size(),
                          s->mymesh.tri.data(), (int)s->mymesh.tri.size() / 3);
}

// render rows [y0,y1) -- Integrator.DoPass, integrators.adb:25-71, for `passes` consecutive passes
uint64_t render_rows(const ao_scene* scn, const ao_params* p, float* accum3, int y0, int y1) {
  PathCtx pc; pc.cx.scn = scn; pc.cx.mesh_mode = p->mesh_mode; pc.cx.rays = 0;
  pc.seed = p->seed; pc.max_depth = p->max_depth; pc.render_type = p->render_type;
  const int W = p->width, H = p->height;
  const float3 bg = f3(p->background);
  for (int y = y0; y < y1; y++) {
    for (int x = 0; x < W; x++) {
      pc.pixel = (uint32_t)(y * W + x);
      float* acc = accum3 + 3 * (size_t)(y * W + x);
      if (p->render_type == AO_RT_DEBUG || p->render_type == AO_RT_WHITTED) {                          // ray_tracer.adb:208-238
        Ray r = make_eye_ray(*scn, W, H, x, y, -1);
        Hit h = Find_Closest_Hit(pc.cx, r);
        float3 c{0, 0, 0};
        if (h.is_hit) c = PALETTE[h.matId % 8];
        acc[0] = c.x; acc[1] = c.y; acc[2] = c.z;
        continue;
      }
      for (int pass = 0; pass < p->passes; pass++) {
        float3 color;
        if (p->anti_aliasing) {
          color = bg;
          for (int i = 0; i < 4; i++) {
            Ray r = make_eye_ray(*scn, W, H, x, y, i);
            pc.sample = (uint32_t)(p->first_sample + pass * 4 + i);
            color = color + PathTrace(pc, r, StartSample, p->max_depth);
          }
        } else {
          Ray r = make_eye_ray(*scn, W, H, x, y, -1);
          pc.sample = (uint32_t)(p->first_sample + pass);
          color = PathTrace(pc, r, StartSample, p->max_depth);
        }
        float3 old{acc[0], acc[1], acc[2]};
        float3 nw = color + old;                                                                        // :51, :63
        acc[0] = nw.x; acc[1] = nw.y; acc[2] = nw.z;
      }
    }
  }
  (void)H;
  return pc.cx.rays;
}

inline uint32_t to_u32_round(float v) {            // Ada float -> Unsigned_32: nearest, ties away from zero
  float f = floorf(v);
  float d = v - f;                                 // exact
  if (d >= 0.5f) f = f + 1.0f;
  if (!(f > 0.0f)) return 0u;                      // ADA-RAISES Constraint_Error for negatives
  return (uint32_t)f;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================

extern "C" {

ao_scene* ao_scene_new(void) {
  ao_scene* s = new ao_scene();
  memset(&s->box, 0, sizeof(s->box)); memset(&s->g_light, 0, sizeof(s->g_light));
  memset(&s->light, 0, sizeof(s->light)); memset(&s->cam, 0, sizeof(s->cam));
  s->light.shape = AO_LIGHT_SPHERE;
  float4x4 I = identity4(); memcpy(s->cam.matrix, &I, sizeof(I));
  return s;
}

void ao_scene_free(ao_scene* s) { if (s) { if (s->bvh) ao_bvh_free(s->bvh); delete s; } }

void ao_set_spheres(ao_scene* s, const ao_sphere* p, int n) { s->spheres.assign(p, p + n); }
void ao_set_cornell(ao_scene* s, const ao_cornell* b) { s->box = *b; }
void ao_set_flatlight(ao_scene* s, const ao_flatlight* l) { s->g_light = *l; }
void ao_set_materials(ao_scene* s, const ao_material* m, int n) { s->materials.assign(m, m + n); }
void ao_set_light(ao_scene* s, const ao_light* l) { s->light = *l; }
void ao_set_camera(ao_scene* s, const ao_camera* c) { s->cam = *c; }

static void fold_bbox(Mesh& m) {                                                                       // geometry.adb:593-607 (positions already transformed)
  for (size_t i = 0; i < m.vert_positions.size(); i++) {
    const float3& v = m.vert_positions[i];
    m.bbox.min.x = min2(m.bbox.min.x, v.x); m.bbox.min.y = min2(m.bbox.min.y, v.y); m.bbox.min.z = min2(m.bbox.min.z, v.z);
    m.bbox.max.x = max2(m.bbox.max.x, v.x); m.bbox.max.y = max2(m.bbox.max.y, v.y); m.bbox.max.z = max2(m.bbox.max.z, v.z);
  }
}

void ao_set_mesh(ao_scene* s, const float* pos3, const float* nrm3, int nverts, const int32_t* tris3, int ntris) {
  Mesh& m = s->mymesh;
  m = Mesh();
  if (s->bvh) { ao_bvh_free(s->bvh); s->bvh = nullptr; }
  m.vert_positions.resize(nverts); m.vert_normals.resize(nverts);
  m.tex_x.assign(nverts, 0.0f); m.tex_y.assign(nverts, 0.0f);
  for (int i = 0; i < nverts; i++) { m.vert_positions[i] = f3(pos3 + 3 * i); m.vert_normals[i] = f3(nrm3 + 3 * i); }
  m.tri.assign(tris3, tris3 + 3 * (size_t)ntris);
  m.material_ids.assign(ntris, 2);
  fold_bbox(m);
}

void ao_set_mesh_tri_materials(ao_scene* s, const int32_t* mats, int ntris) {
  Mesh& m = s->mymesh;
  if ((size_t)ntris * 3 != m.tri.size()) return;
  m.material_ids.assign(mats, mats + ntris);
  m.custom_material_ids = true;
}

int ao_load_vsgf(ao_scene* s, const char* path, const float* matrix16) {                               // geometry.adb:499-609
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  struct { int64_t fileSizeInBytes; int32_t verticesNum, indicesNum, materialsNum, flags; } header;
  if (fread(&header, 24, 1, f) != 1) { fclose(f); return -2; }
  int vn = header.verticesNum, in = header.indicesNum;
  if (vn <= 0 || in < 0) { fclose(f); return -3; }
  Mesh m;
  m.vert_positions.resize(vn); m.vert_normals.resize(vn); m.tex_x.assign(vn, 0.0f); m.tex_y.assign(vn, 0.0f);
  m.tri.resize(3 * (size_t)(in / 3)); m.material_ids.resize(in / 3);
  std::vector<float> tmp(4 * (size_t)vn);
  bool ok = fread(tmp.data(), 16, vn, f) == (size_t)vn;
  for (int i = 0; ok && i < vn; i++) m.vert_positions[i] = {tmp[4 * i], tmp[4 * i + 1], tmp[4 * i + 2]};
  ok = ok && fread(tmp.data(), 16, vn, f) == (size_t)vn;
  for (int i = 0; ok && i < vn; i++) m.vert_normals[i] = {tmp[4 * i], tmp[4 * i + 1], tmp[4 * i + 2]};
  ok = ok && fread(tmp.data(), 8, vn, f) == (size_t)vn;                      // texcoords read and zeroed (:562-567)
  if (ok && header.flags != 0) ok = fread(tmp.data(), 16, vn, f) == (size_t)vn;   // tangents skipped (:571-575)
  ok = ok && fread(m.tri.data(), 12, in / 3, f) == (size_t)(in / 3);
  ok = ok && fread(m.material_ids.data(), 4, in / 3, f) == (size_t)(in / 3);
  fclose(f);
  if (!ok) return -4;
  float4x4 M; memcpy(&M, matrix16, sizeof(M));
  for (int i = 0; i < vn; i++) m.vert_positions[i] = mul(M, m.vert_positions[i]);                      // :595
  m.bbox = s->mymesh.bbox;               // "never reset": keeps whatever the record held (zeros on a fresh scene)
  fold_bbox(m);
  if (s->bvh) { ao_bvh_free(s->bvh); s->bvh = nullptr; }
  s->mymesh = m;
  return 0;
}

int ao_scene_init_cornell(ao_scene* s, const char* vsgf_path) {                                        // scene.adb:89-217
  const float intensity[3] = {20.0f, 20.0f, 20.0f};
  // lights (:102-148)
  memset(&s->light, 0, sizeof(s->light));
  s->light.shape = AO_LIGHT_SPHERE;                                  // a_scn.g_lightRef := sphLight (:128)
  s->light.boxMin[0] = -0.75f; s->light.boxMin[1] = 4.98f; s->light.boxMin[2] = 1.25f;
  s->light.boxMax[0] = 0.75f;  s->light.boxMax[1] = 4.98f; s->light.boxMax[2] = 3.25f;
  s->light.normal[0] = 0; s->light.normal[1] = -1.0f; s->light.normal[2] = 0;
  s->light.center[0] = 0.0f; s->light.center[1] = 4.5f; s->light.center[2] = 1.0f;
  s->light.radius = 0.5f;
  for (int i = 0; i < 3; i++) s->light.intensity[i] = 0.5f * intensity[i];                            // intensity*0.5 (:121)
  s->light.surfaceArea = 4.0f * M_PI_F * s->light.radius * s->light.radius;                            // :122
  memset(&s->g_light, 0, sizeof(s->g_light));
  // spheres (:139-144, 182-192)
  s->spheres.resize(3);
  s->spheres[0] = {{-1.5f, 1.0f, 1.5f}, 1.0f, 8};                   // smatPhong
  s->spheres[1] = {{1.4f, 1.0f, 3.0f}, 1.0f, 0};                    // smatGlass
  s->spheres[2] = {{0.0f, 4.5f, 1.0f}, 0.5f, 4};                    // MaterialLight(lref => g_lightRef)
  // materials (:153-185)
  s->materials.assign(11, ao_material{AO_MAT_NULL, {0, 0, 0}, {0, 0, 0}, 0.0f, -1});
  s->materials[0]  = {AO_MAT_GLASS,   {0.75f, 0.75f, 0.75f}, {0.85f, 0.85f, 0.85f}, 1.75f, -1};
  s->materials[1]  = {AO_MAT_LAMBERT, {0.5f, 0.5f, 0.5f},    {0, 0, 0}, 0.0f, -1};
  s->materials[2]  = {AO_MAT_LAMBERT, {0.25f, 0.5f, 0.0f},   {0, 0, 0}, 0.0f, -1};
  s->materials[3]  = {AO_MAT_LAMBERT, {0.5f, 0.0f, 0.0f},    {0, 0, 0}, 0.0f, -1};
  s->materials[4]  = {AO_MAT_LIGHT,   {0, 0, 0},             {0, 0, 0}, 0.0f, 0};
  s->materials[5]  = {AO_MAT_MIRROR,  {0.75f, 0.75f, 0.75f}, {0, 0, 0}, 0.0f, -1};
  s->materials[8]  = {AO_MAT_PHONG,   {0.75f, 0.75f, 0.75f}, {0, 0, 0}, 80.0f, -1};
  s->materials[9]  = s->materials[1];
  s->materials[10] = s->materials[1];
  // cornell box (scene.ads:75-80)
  const ao_cornell box = {{-2.5f, 0.0f, 0.0f}, {2.5f, 5.0f, 5.0f}, {2, 3, 1, 1, 8, 1},
                          {1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1}};
  s->box = box;
  // mesh (:194-208)
  float4x4 mrot = RotationMatrix(-M_PI_F / 6.0f, {0.0f, 1.0f, 0.0f});
  float4x4 mscale = identity4(), mtans = identity4();
  mtans.m[0][3] = -0.75f; mtans.m[1][3] = 0.1f; mtans.m[2][3] = 3.1f; mtans.m[3][3] = 1.0f;            // SetCol(mtans, 3, ...)
  mscale.m[0][0] = 2.0f; mscale.m[1][1] = 2.0f; mscale.m[2][2] = 2.0f;
  float4x4 M = matmul(matmul(mtans, mrot), mscale);
  s->mymesh = Mesh();
  int rc = ao_load_vsgf(s, vsgf_path, &M.m[0][0]);
  // camera (:212-215)
  memset(&s->cam, 0, sizeof(s->cam));
  s->cam.pos[0] = 0.0f; s->cam.pos[1] = 2.55f; s->cam.pos[2] = 12.5f;
  s->cam.up[1] = 1.0f;
  float4x4 I = identity4(); memcpy(s->cam.matrix, &I, sizeof(I));
  return rc;
}

int ao_get_counts(const ao_scene* s, int* nspheres, int* nmaterials, int* nverts, int* ntris) {
  if (nspheres) *nspheres = (int)s->spheres.size();
  if (nmaterials) *nmaterials = (int)s->materials.size();
  if (nverts) *nverts = (int)s->mymesh.vert_positions.size();
  if (ntris) *ntris = (int)s->mymesh.tri.size() / 3;
  return 0;
}
void ao_get_spheres(const ao_scene* s, ao_sphere* out) { if (!s->spheres.empty()) memcpy(out, s->spheres.data(), s->spheres.size() * sizeof(ao_sphere)); }
void ao_get_cornell(const ao_scene* s, ao_cornell* out) { *out = s->box; }
void ao_get_materials(const ao_scene* s, ao_material* out) { if (!s->materials.empty()) memcpy(out, s->materials.data(), s->materials.size() * sizeof(ao_material)); }
void ao_get_light(const ao_scene* s, ao_light* out) { *out = s->light; }
void ao_get_camera(const ao_scene* s, ao_camera* out) { *out = s->cam; }
void ao_get_mesh(const ao_scene* s, float* pos3, float* nrm3, int32_t* tris3, float* bbox6) {
  const Mesh& m = s->mymesh;
  for (size_t i = 0; i < m.vert_positions.size(); i++) {
    if (pos3) { pos3[3 * i] = m.vert_positions[i].x; pos3[3 * i + 1] = m.vert_positions[i].y; pos3[3 * i + 2] = m.vert_positions[i].z; }
    if (nrm3) { nrm3[3 * i] = m.vert_normals[i].x; nrm3[3 * i + 1] = m.vert_normals[i].y; nrm3[3 * i + 2] = m.vert_normals[i].z; }
  }
  if (tris3 && !m.tri.empty()) memcpy(tris3, m.tri.data(), m.tri.size() * sizeof(int32_t));
  if (bbox6) { bbox6[0] = m.bbox.min.x; bbox6[1] = m.bbox.min.y; bbox6[2] = m.bbox.min.z; bbox6[3] = m.bbox.max.x; bbox6[4] = m.bbox.max.y; bbox6[5] = m.bbox.max.z; }
}

void ao_closest_hit_batch(const ao_scene* s, int mesh_mode, int n, const ao_ray* rays, ao_hit* hits) {
  ensure_bvh(s, mesh_mode);
  Ctx cx; cx.scn = s; cx.mesh_mode = mesh_mode;
  for (int i = 0; i < n; i++) {
    Ray r; r.origin = f3(rays[i].origin); r.direction = f3(rays[i].direction);
    hit_to_c(Find_Closest_Hit(cx, r), &hits[i]);
  }
}

void ao_eye_ray(const ao_scene* s, int width, int height, int x, int y, int sub, ao_ray* out) {
  Ray r = make_eye_ray(*s, width, height, x, y, sub);
  out->origin[0] = r.origin.x; out->origin[1] = r.origin.y; out->origin[2] = r.origin.z;
  out->direction[0] = r.direction.x; out->direction[1] = r.direction.y; out->direction[2] = r.direction.z;
}

void ao_primary_hits(const ao_scene* s, const ao_params* p, ao_hit* hits) {
  ensure_bvh(s, p->mesh_mode);
  Ctx cx; cx.scn = s; cx.mesh_mode = p->mesh_mode;
  for (int y = 0; y < p->height; y++)
    for (int x = 0; x < p->width; x++) {
      Ray r = make_eye_ray(*s, p->width, p->height, x, y, -1);
      hit_to_c(Find_Closest_Hit(cx, r), &hits[(size_t)y * p->width + x]);
    }
}

int ao_compute_shadow(const ao_scene* s, int mesh_mode, const float hit_pos[3], const float lpos[3]) {
  ensure_bvh(s, mesh_mode);
  Ctx cx; cx.scn = s; cx.mesh_mode = mesh_mode;
  return Compute_Shadow(cx, f3(hit_pos), f3(lpos)) ? 1 : 0;
}

uint64_t ao_render(const ao_scene* s, const ao_params* p, float* accum3) {
  ensure_bvh(s, p->mesh_mode);
  int y0 = p->y0, y1 = p->y1;
  if (y0 == 0 && y1 == 0) y1 = p->height;
  int T = p->threads > 0 ? p->threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > y1 - y0) T = std::max(1, y1 - y0);
  std::vector<uint64_t> rays(T, 0);
  std::atomic<int> next_row{y0};
  auto worker = [&](int tid) {
    for (;;) {                       // dynamic row striping: no two workers touch the same pixel, no lock needed
      int y = next_row.fetch_add(1);
      if (y >= y1) break;
      rays[tid] += render_rows(s, p, accum3, y, y + 1);
    }
  };
  if (T == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  uint64_t total = 0; for (uint64_t r : rays) total += r;
  return total;
}

// The reference's own parallel scheme, for TIMING ONLY (mode B): `tasks` workers, each renders the WHOLE
// image for one pass with its own sample indices (task k = DoPass unit k, like Path_Trace_Thread k of ray_tracer.adb:142-194) and adds
// every pixel into the shared accumulator under one process-wide lock (GNAT.Task_Lock, ray_tracer-integrators.adb:50-52).  The sum per
// pixel is the same set of terms as ao_render with passes = tasks, added in a scheduling-dependent order.
uint64_t ao_render_task_lock(const ao_scene* s, const ao_params* p, int tasks, float* accum3) {
  ensure_bvh(s, p->mesh_mode);
  if (tasks < 1) tasks = 1;
  const int W = p->width, H = p->height;
  static std::mutex task_lock;
  std::vector<uint64_t> rays(tasks, 0);
  auto worker = [&](int k) {
    PathCtx pc; pc.cx.scn = s; pc.cx.mesh_mode = p->mesh_mode; pc.cx.rays = 0;
    pc.seed = p->seed; pc.max_depth = p->max_depth; pc.render_type = p->render_type;
    const float3 bg = f3(p->background);
    for (int y = 0; y < H; y++) {
      for (int x = 0; x < W; x++) {
        pc.pixel = (uint32_t)(y * W + x);
        float3 color;
        if (p->anti_aliasing) {
          color = bg;
          for (int i = 0; i < 4; i++) {
            Ray r = make_eye_ray(*s, W, H, x, y, i);
            pc.sample = (uint32_t)(p->first_sample + k * 4 + i);
            color = color + PathTrace(pc, r, StartSample, p->max_depth);
          }
        } else {
          Ray r = make_eye_ray(*s, W, H, x, y, -1);
          pc.sample = (uint32_t)(p->first_sample + k);
          color = PathTrace(pc, r, StartSample, p->max_depth);
        }
        float* acc = accum3 + 3 * (size_t)(y * W + x);
        std::lock_guard<std::mutex> lk(task_lock);                                                      // :50-52
        acc[0] = color.x + acc[0]; acc[1] = color.y + acc[1]; acc[2] = color.z + acc[2];
      }
    }
    rays[k] = pc.cx.rays;
  };
  std::vector<std::thread> th;
  for (int k = 0; k < tasks; k++) th.emplace_back(worker, k);
  for (auto& t : th) t.join();
  uint64_t total = 0; for (uint64_t r : rays) total += r;
  return total;
}

void ao_resolve(const float* accum3, int npixels, int spp, uint32_t* ldr) {                            // ray_tracer.adb:281-291, 19-57
  float normC = 1.0f / (float)spp;
  for (int i = 0; i < npixels; i++) {
    float3 rgb = f3(accum3 + 3 * (size_t)i) * normC;
    rgb.x = sqrtf(rgb.x); rgb.y = sqrtf(rgb.y); rgb.z = sqrtf(rgb.z);   // ** (1.0/g_gamma), g_gamma = 2.0 -> Sqrt
    float cr = min2(rgb.x, 1.0f), cg = min2(rgb.y, 1.0f), cb = min2(rgb.z, 1.0f);                       // ToneMapping
    uint32_t red = to_u32_round(cr * 255.0f), green = to_u32_round(cg * 255.0f), blue = to_u32_round(cb * 255.0f);
    ldr[i] = red | (green << 8) | (blue << 16);                                                         // ColorToUnsigned_32
  }
}

int ao_save_bmp(const char* path, const uint32_t* data, int width, int height) {                        // bitmap.adb:31-85
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  unsigned char hdr[54]; memset(hdr, 0, sizeof(hdr));
  auto put16 = [&](int o, uint32_t v) { hdr[o] = v & 255; hdr[o + 1] = (v >> 8) & 255; };
  auto put32 = [&](int o, uint32_t v) { put16(o, v & 0xffff); put16(o + 2, v >> 16); };
  put16(0, 0x4d42); put32(2, 14 + 40 + (uint32_t)(width * height * 3)); put32(10, 14 + 40);
  put32(14, 40); put32(18, (uint32_t)width); put32(22, (uint32_t)height); put16(26, 1); put16(28, 24);
  fwrite(hdr, 1, 54, f);
  std::vector<unsigned char> row(3 * (size_t)width);
  for (int y = 0; y < height; y++) {            // data(0) first: row y=0 is the bottom row of the BMP; no row padding
    for (int x = 0; x < width; x++) {
      uint32_t px = data[(size_t)y * width + x];
      row[3 * x] = (px >> 16) & 255; row[3 * x + 1] = (px >> 8) & 255; row[3 * x + 2] = px & 255;       // Pixel'Write order r,g,b fields = (bits 16..23, 8..15, 0..7)
    }
    fwrite(row.data(), 1, row.size(), f);
  }
  fclose(f);
  return 0;
}

void ao_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
void ao_sincos(float x, float* s, float* c) { det_sincos(x, s, c); }
float ao_pow(float x, float y) { return pow_ada(x, y); }
void ao_use_libm(int on) { g_use_libm.store(on ? 1 : 0); }

ao_bvh* ao_bvh_build(const float* pos3, int nverts, const int32_t* tris3, int ntris) {
  ao_bvh* b = new ao_bvh();
  b->pos_own.assign(pos3, pos3 + 3 * (size_t)nverts);
  b->tris_own.assign(tris3, tris3 + 3 * (size_t)ntris);
  b->pos = b->pos_own.data(); b->tris = b->tris_own.data(); b->ntris = ntris;
  b->order.resize(ntris);
  std::vector<float> tb(6 * (size_t)ntris);
  for (int t = 0; t < ntris; t++) {
    b->order[t] = t;
    tri_bounds(b->pos, b->tris + 3 * t, &tb[6 * (size_t)t], &tb[6 * (size_t)t + 3]);
  }
  if (ntris > 0) {
    b->nodes.resize(2 * (size_t)ntris);
    std::atomic<int> next_node(1);
    bvh_build_rec(*b, tb, next_node, 0, 0, ntris, 0);
    b->nodes.resize(next_node.load());
  }
  return b;
}

void ao_bvh_free(ao_bvh* b) { delete b; }

void ao_bvh_closest(const ao_bvh* b, int n, const ao_ray* rays, float tmin, float tmax,
                    int32_t* out_prim, float* out_t, float* out_u, float* out_v, int threads) {
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      Ray r; r.origin = f3(rays[i].origin); r.direction = f3(rays[i].direction);
      int prim; float t, u, v;
      bvh_closest_one(*b, r, tmin, tmax, &prim, &t, &u, &v);
      out_prim[i] = prim; if (out_t) out_t[i] = t; if (out_u) out_u[i] = u; if (out_v) out_v[i] = v;
    }
  };
  if (T == 1 || n < 1024) { work(0, n); return; }
  std::vector<std::thread> th;
  int chunk = (n + T - 1) / T;
  for (int t = 0; t < T; t++) { int lo = t * chunk, hi = std::min(n, lo + chunk); if (lo < hi) th.emplace_back(work, lo, hi); }
  for (auto& t : th) t.join();
}

void ao_bruteforce_closest(const float* pos3, const int32_t* tris3, int ntris, int n, const ao_ray* rays, float tmin, float tmax,
                           int32_t* out_prim, float* out_t, float* out_u, float* out_v) {
  for (int i = 0; i < n; i++) {
    Ray r; r.origin = f3(rays[i].origin); r.direction = f3(rays[i].direction);
    int prim; float t, u, v;
    bruteforce_closest_one(pos3, tris3, ntris, r, tmin, tmax, &prim, &t, &u, &v);
    out_prim[i] = prim; if (out_t) out_t[i] = t; if (out_u) out_u[i] = u; if (out_v) out_v[i] = v;
  }
}

// ---- batched access to the shading pieces, for the independent numpy restatement of tests/ada_numpy_shading.py -----------------
void ao_sincos_batch(int n, const float* x, float* s, float* c) { for (int i = 0; i < n; i++) sincos_ada(x[i], &s[i], &c[i]); }
void ao_expon_batch(int n, const float* left, const float* right, float* out) { for (int i = 0; i < n; i++) out[i] = ada_expon(left[i], right[i]); }
/* Lights.Sample (lights.adb:56-79, 197-240) for n illuminated points; u = (u1, u2) per point */
void ao_light_sample_batch(const ao_scene* s, int n, const float* u2, const float* p3, float* out_pos3, float* out_pdf) {
  for (int i = 0; i < n; i++) {
    BounceRand br; br.u[0] = u2[2 * i]; br.u[1] = u2[2 * i + 1]; br.u[2] = 0.0f; br.u[3] = 0.0f;
    ShadowSample ss = Light_Sample(s->light, br, f3(p3 + 3 * i));
    out_pos3[3 * i] = ss.pos.x; out_pos3[3 * i + 1] = ss.pos.y; out_pos3[3 * i + 2] = ss.pos.z; out_pdf[i] = ss.pdf;
  }
}
/* Lights.EvalPDF (lights.adb:81-85, 242-255) */
void ao_light_eval_pdf_batch(const ao_scene* s, int n, const float* p3, const float* dir3, const float* t, float* out) {
  for (int i = 0; i < n; i++) out[i] = Light_EvalPDF(s->light, f3(p3 + 3 * i), f3(dir3 + 3 * i), t[i]);
}
/* SampleAndEvalBxDF of material mat_id (materials.adb:197-215, 247-254, 295-331, 363-387); r = (r1, r2) per sample */
void ao_mat_sample_batch(const ao_scene* s, int mat_id, int n, const float* r2, const float* ray_dir3, const float* normal3,
                         float* out_color3, float* out_dir3, float* out_pdf, int32_t* out_spec) {
  const ao_material& m = s->materials[mat_id];
  for (int i = 0; i < n; i++) {
    BounceRand br; br.u[0] = 0.0f; br.u[1] = 0.0f; br.u[2] = r2[2 * i]; br.u[3] = r2[2 * i + 1];
    MatSample ms = Mat_SampleAndEvalBxDF(m, br, f3(ray_dir3 + 3 * i), f3(normal3 + 3 * i));
    out_color3[3 * i] = ms.color.x; out_color3[3 * i + 1] = ms.color.y; out_color3[3 * i + 2] = ms.color.z;
    out_dir3[3 * i] = ms.direction.x; out_dir3[3 * i + 1] = ms.direction.y; out_dir3[3 * i + 2] = ms.direction.z;
    out_pdf[i] = ms.pdf; out_spec[i] = ms.pureSpecular ? 1 : 0;
  }
}
/* EvalBxDF / EvalPDF of material mat_id for light direction l, view direction v, normal n (materials.adb:217-226, 389-410, ...) */
void ao_mat_eval_batch(const ao_scene* s, int mat_id, int n, const float* l3, const float* v3, const float* n3, float* out_bxdf3, float* out_pdf) {
  const ao_material& m = s->materials[mat_id];
  for (int i = 0; i < n; i++) {
    float3 b = Mat_EvalBxDF(m, f3(l3 + 3 * i), f3(v3 + 3 * i), f3(n3 + 3 * i));
    out_bxdf3[3 * i] = b.x; out_bxdf3[3 * i + 1] = b.y; out_bxdf3[3 * i + 2] = b.z;
    out_pdf[i] = Mat_EvalPDF(m, f3(l3 + 3 * i), f3(v3 + 3 * i), f3(n3 + 3 * i));
  }
}

// the same query with Embree's two-sided triangle test (det of either sign): the selectable behaviour of the legacy seam
void ao_bruteforce_closest_two_sided(const float* pos3, const int32_t* tris3, int ntris, int n, const ao_ray* rays, float tmin, float tmax,
                                     int32_t* out_prim, float* out_t, float* out_u, float* out_v) {
  for (int i = 0; i < n; i++) {
    Ray r; r.origin = f3(rays[i].origin); r.direction = f3(rays[i].direction);
    int best = -1; float best_t = tmax, bu = 0, bv = 0;
    for (int t = 0; t < ntris; t++) {
      const int32_t* tr = tris3 + 3 * t;
      float3 A = f3(pos3 + 3 * tr[0]), B = f3(pos3 + 3 * tr[1]), C = f3(pos3 + 3 * tr[2]);
      float3 edge1 = B - A, edge2 = C - A;
      float3 pvec = cross(r.direction, edge2);
      float3 tvec = r.origin - A;
      float det = dot(edge1, pvec);
      if (det == 0.0f) continue;
      float invDet = 1.0f / det;
      float v = dot(tvec, pvec) * invDet;
      float3 qvec = cross(tvec, edge1);
      float u = dot(qvec, r.direction) * invDet;
      float tt = dot(edge2, qvec) * invDet;
      if (v > 0.0f && u > 0.0f && u + v < 1.0f && tt > tmin && tt < best_t) { best = t; best_t = tt; bu = u; bv = v; }
    }
    out_prim[i] = best; if (out_t) out_t[i] = best_t; if (out_u) out_u[i] = bu; if (out_v) out_v[i] = bv;
  }
}

} // extern "C"
