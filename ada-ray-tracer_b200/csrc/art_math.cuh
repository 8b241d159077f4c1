// art_math.cuh -- binary32 vector math, deterministic elementary functions and the counter-based RNG
// used by the render kernels.
//
// Numerics contract (must hold for every translation unit that includes this file): compiled with
//   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// so that each + - * / sqrt is one correctly rounded IEEE binary32 operation, evaluated in the order
// written.  The expressions follow the reference's Ada sources operation for operation
// (generic_vector_math.adb, vector_math.adb); the reference's Ada.Numerics run-time pieces (Float_Random,
// sin/cos/"**") are replaced by Philox4x32-10 and by the fixed-sequence kernels below.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "art_types.cuh"

#define ART_D __device__ __forceinline__

namespace art {


#define ART_FLT_LAST 3.402823466e+38f                    // vector_math.ads:17  infinity = Float'Last
#define ART_PI       3.14159265358979323846f              // vector_math.ads:19
#define ART_INV_PI   0.318309886183790671538f             // vector_math.ads:20

// generic_vector_math.adb:19-63 -- comparisons spelled out (NaN ordering matters in the slab test)
ART_D float min2(float a, float b) { return (a < b) ? a : b; }
ART_D float max2(float a, float b) { return (a >= b) ? a : b; }
ART_D float max3(float a, float b, float c) {
  if (a >= b && a >= c) return a;
  else if (b >= c && b >= a) return b;
  else return c;
}
ART_D float clampf(float x, float a, float b) { return min2(max2(x, a), b); }

ART_D f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
ART_D f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
ART_D f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }     // :81-88
ART_D f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }     // :92-99
ART_D f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }     // :101-108
ART_D f3 operator*(f3 a, float k) { return mk3(k * a.x, k * a.y, k * a.z); }        // :148-151
ART_D f3 operator*(float k, f3 a) { return mk3(k * a.x, k * a.y, k * a.z); }        // :153-156
ART_D float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }           // :110-113
ART_D f3 cross(f3 a, f3 b) {                                                          // :115-122
  return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
ART_D f3 neg(f3 a) { return mk3(-a.x, -a.y, -a.z); }                                  // (-1.0)*v, exact

ART_D float sign_ada(float x) { return (x >= 0.0f) ? 1.0f : -1.0f; }                 // vector_math.adb:49-56
ART_D float lerp_ada(float t, float a, float b) { return (1.0f - t) * a + t * b; }   // :59-62
ART_D f3 normalize(f3 a) {                                                            // :64-72
  float l_inv = 1.0f / sqrtf(dot(a, a));
  return mk3(l_inv * a.x, l_inv * a.y, l_inv * a.z);
}
ART_D float length(f3 a) { return sqrtf(dot(a, a)); }                                 // :74-77
ART_D f3 reflect(f3 dir, f3 normal) {                                                 // :79-82
  float d = dot(dir, normal);
  return normalize(mk3((-2.0f) * (d * normal.x) + dir.x, (-2.0f) * (d * normal.y) + dir.y, (-2.0f) * (d * normal.z) + dir.z));
}

// ------------------------------------------------------------------------------------------------
// deterministic elementary functions: fixed operation sequences in binary32
// ------------------------------------------------------------------------------------------------

// sin/cos: Cody-Waite reduction by pi/2 (three-part constant), degree-7/8 minimax polynomials
ART_D void det_sincos(float x, float& s_out, float& c_out) {
  float q  = x * 0.636619772367581343f;
  float qf = floorf(q + 0.5f);
  int   k  = (int)qf;
  float r  = x - qf * 1.5703125f;
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
  int kk = k & 3;
  float ss = (kk & 1) ? c : s;
  float cc = (kk & 1) ? s : c;
  s_out = (kk & 2) ? -ss : ss;
  c_out = ((kk + 1) & 2) ? -cc : cc;
}

ART_D float det_log2(float x) {
  uint32_t ix = __float_as_uint(x);
  int e = 0;
  if (ix < 0x00800000u) { x = x * 8388608.0f; ix = __float_as_uint(x); e = -23; }
  e += (int)(ix >> 23) - 127;
  float m = __uint_as_float((ix & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }
  float s  = (m - 1.0f) / (m + 1.0f);
  float s2 = s * s;
  float p  = 0.111111111111f;
  p = p * s2 + 0.142857142857f;
  p = p * s2 + 0.2f;
  p = p * s2 + 0.333333333333f;
  p = p * s2 * s + s;
  float lnm = p + p;
  return (float)e + lnm * 1.44269504088896341f;
}

ART_D float det_exp2(float t) {
  if (!(t > -150.0f)) { if (t != t) return t; return 0.0f; }
  if (t > 128.0f) return __int_as_float(0x7f800000);
  float kf = floorf(t + 0.5f);
  int   k  = (int)kf;
  float z  = (t - kf) * 0.693147180559945309f;
  float p  = 1.98412698413e-4f;
  p = p * z + 1.38888888889e-3f;
  p = p * z + 8.33333333333e-3f;
  p = p * z + 4.16666666667e-2f;
  p = p * z + 1.66666666667e-1f;
  p = p * z + 0.5f;
  p = p * z + 1.0f;
  p = p * z + 1.0f;
  int k1 = k / 2, k2 = k - k1;
  float s1 = __uint_as_float((uint32_t)(k1 + 127) << 23);
  float s2 = __uint_as_float((uint32_t)(k2 + 127) << 23);
  return p * s1 * s2;
}

ART_D float det_powi(float x, int n) {
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

// Left ** Right (Ada.Numerics.Generic_Elementary_Functions semantics, see oracle notes in DESIGN.md)
ART_D float ada_expon(float left, float right) {
  if (left < 0.0f) return 0.0f;
  if (left == 0.0f && right == 0.0f) return 1.0f;
  if (right == 0.0f) return 1.0f;
  if (left == 0.0f) { if (right < 0.0f) return __int_as_float(0x7f800000); return 0.0f; }
  if (left == 1.0f) return 1.0f;
  if (right == 1.0f) return left;
  if (right == 2.0f) return left * left;
  if (right == 0.5f) return sqrtf(left);
  float a_right = fabsf(right);
  if (a_right > 1.0f && a_right < 2147483648.0f) {
    int   int_part = (int)a_right;
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
    return (right >= 0.0f) ? result : 1.0f / result;
  }
  return det_exp2(right * det_log2(left));
}

// function pow, vector_math.adb:24-47
ART_D float pow_ada(float left, float right) {
  if (left == 0.0f && right == 0.0f) return 1.0f;
  else if (left < 0.0f) return 0.0f;
  else if (right == 0.0f) return 1.0f;
  else if (left == 0.0f) { return (right < 0.0f) ? __int_as_float(0x7f800000) : 0.0f; }
  else if (left == 1.0f) return 1.0f;
  else if (right == 1.0f) return left;
  else return ada_expon(left, right);
}

ART_D float sqrt_clamped(float x) { return sqrtf(max2(x, 0.0f)); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10: key = (seed, pixel), counter = (sample, bounce, 0, 0) -> 4 uniforms in [0,1)
// ------------------------------------------------------------------------------------------------

struct BounceRand { float u0, u1, u2, u3; };   // light u1,u2 ; bsdf r1,r2

ART_D BounceRand bounce_rand(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce) {
  uint32_t c0 = sample, c1 = bounce, c2 = 0u, c3 = 0u, k0 = seed, k1 = pixel;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  BounceRand b;
  b.u0 = (float)(c0 >> 8) * (1.0f / 16777216.0f);
  b.u1 = (float)(c1 >> 8) * (1.0f / 16777216.0f);
  b.u2 = (float)(c2 >> 8) * (1.0f / 16777216.0f);
  b.u3 = (float)(c3 >> 8) * (1.0f / 16777216.0f);
  return b;
}

}  // namespace art
