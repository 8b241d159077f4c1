// lbvh_build.cu -- GPU LBVH build (K3): replaces the Embree BVH builder invoked by rtcCommitScene
// (cpp/embree_connect.cpp:136,170,243).
//
//   1. tri_bounds_kernel    per-triangle AABB + centroid, scene bounds by warp reduce + atomics
//   2. morton_kernel        63-bit Morton code of the centroid (21 bits per axis)
//   3. cub::DeviceRadixSort (key = Morton code, value = triangle id)
//   4. karras_kernel        binary radix tree (Karras 2012), ties broken by the sorted position
//   5. refit_kernel         bottom-up AABB union with one atomic flag per internal node
//   6. pack_kernel          32-byte nodes (both child boxes quantised to 8 bits inside the node's box + child link), 48-byte triangle slots (A|id, e1, e2)
//
// Bit-exactness note: e1 = B - A and e2 = C - A are the same binary32 subtractions the reference performs per ray
// (geometry.adb:238-239); they are done once here.  Boxes are only a culling device and are padded outward.
#include "gcore_internal.h"

#include <cub/cub.cuh>
#include <cfloat>

namespace {

struct Aabb { float mn[3], mx[3]; };

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  // valid for any sign: ordered-int trick
  if (v >= 0.0f) atomicMin((int*)addr, __float_as_int(v));
  else atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax((int*)addr, __float_as_int(v));
  else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__global__ void init_bounds_kernel(float* scene_bounds) {
  if (threadIdx.x < 3) scene_bounds[threadIdx.x] = FLT_MAX;
  else if (threadIdx.x < 6) scene_bounds[threadIdx.x] = -FLT_MAX;
}

__global__ void tri_bounds_kernel(const float* __restrict__ verts, const int* __restrict__ tris, int ntris,
                                  Aabb* __restrict__ boxes, float* __restrict__ scene_bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < ntris) {
    int ia = tris[3 * i], ib = tris[3 * i + 1], ic = tris[3 * i + 2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float a = verts[3 * (size_t)ia + k], b = verts[3 * (size_t)ib + k], c = verts[3 * (size_t)ic + k];
      float lo = fminf(a, fminf(b, c)), hi = fmaxf(a, fmaxf(b, c));
      // outward padding: a few ulps of the coordinate magnitude plus an absolute floor
      float pad = fmaxf(fmaxf(fabsf(lo), fabsf(hi)) * 4.0e-6f, 1.0e-7f);
      mn[k] = lo - pad; mx[k] = hi + pad;
    }
    Aabb bx;
    for (int k = 0; k < 3; k++) { bx.mn[k] = mn[k]; bx.mx[k] = mx[k]; }
    boxes[i] = bx;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float lo = mn[k], hi = mx[k];
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomic_min_float(&scene_bounds[k], lo); atomic_max_float(&scene_bounds[3 + k], hi); }
  }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {   // spread 21 bits to every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8)  & 0x100f00f00f00f00full;
  v = (v | v << 4)  & 0x10c30c30c30c30c3ull;
  v = (v | v << 2)  & 0x1249249249249249ull;
  return v;
}

__global__ void morton_kernel(const Aabb* __restrict__ boxes, int ntris, const float* __restrict__ scene_bounds,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntris) return;
  unsigned long long code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float lo = scene_bounds[k], hi = scene_bounds[3 + k];
    float c = 0.5f * (boxes[i].mn[k] + boxes[i].mx[k]);
    float ext = hi - lo;
    float f = (ext > 0.0f) ? (c - lo) / ext : 0.0f;
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    unsigned long long q = (unsigned long long)(f * 2097151.0f);
    code |= expand21(q) << (2 - k);
  }
  keys[i] = code;
  vals[i] = i;
}

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long x = keys[i] ^ keys[j];
  if (x == 0ull) return 64 + __clz(i ^ j);
  return __clzll((long long)x);
}

// Karras 2012, one thread per internal node.  child < 0 encodes leaf ~slot (slot = sorted position).
__global__ void karras_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                              int* __restrict__ parent_internal, int* __restrict__ parent_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  int lc = (lo == gamma) ? ~gamma : gamma;
  int rc = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
  left[i] = lc; right[i] = rc;
  if (lc < 0) parent_leaf[~lc] = i; else parent_internal[lc] = i;
  if (rc < 0) parent_leaf[~rc] = i; else parent_internal[rc] = i;
  if (i == 0) parent_internal[0] = -1;
}

// child boxes were written by other threads of this launch: read them through L2 (ld.global.cg), not L1
__device__ __forceinline__ Aabb load_box_cg(const Aabb* p) {
  Aabb r;
  const float* f = (const float*)p;
#pragma unroll
  for (int k = 0; k < 3; k++) { r.mn[k] = __ldcg(f + k); r.mx[k] = __ldcg(f + 3 + k); }
  return r;
}

__global__ void refit_kernel(int n, const int* __restrict__ vals, const Aabb* __restrict__ tri_boxes, const int* __restrict__ left,
                             const int* __restrict__ right, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf,
                             Aabb* node_boxes, int* flags) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  int node = parent_leaf[slot];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&flags[node], 1) == 0) return;            // first child to arrive leaves; the second one merges
    __threadfence();
    Aabb a, b;
    int lc = left[node], rc = right[node];
    a = (lc < 0) ? tri_boxes[vals[~lc]] : load_box_cg(&node_boxes[lc]);
    b = (rc < 0) ? tri_boxes[vals[~rc]] : load_box_cg(&node_boxes[rc]);
    Aabb u;
#pragma unroll
    for (int k = 0; k < 3; k++) { u.mn[k] = fminf(a.mn[k], b.mn[k]); u.mx[k] = fmaxf(a.mx[k], b.mx[k]); }
    node_boxes[node] = u;
    node = parent_internal[node];
  }
}

// biased exponent eb such that 255 * 2^(eb - 127) >= extent (with room for the rounding of the decode)
__device__ __forceinline__ unsigned int quant_exponent(float extent) {
  if (!(extent > 0.0f)) return 1u;
  int e;
  frexpf(extent * (1.0f / 255.0f) * 1.0001f, &e);                 // extent/255 = m * 2^e with m in [0.5, 1)  ->  2^e >= extent/255
  int eb = e + 127;
  return (unsigned int)min(max(eb, 1), 254);
}

// 32-byte node: the children's boxes as 8-bit offsets inside the node's own box (origin + one power-of-two step per axis).
//   q0 = (origin.x, origin.y, origin.z, bits: ex | ey << 8 | ez << 16 | leaf flags << 24)        flags: 1 = child 0 is a leaf, 2 = child 1
//   q1 = bits (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x), (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y), (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z), gamma
// children are (gamma, gamma + 1): internal node index or sorted triangle slot (Karras: the split position names both).
// Offsets are rounded outward and checked against the decode the traversal performs, so a decoded box always contains the
// exact one -- boxes only cull, results stay those of the triangle test.
__global__ void pack_nodes_kernel(int n, const int* __restrict__ vals, const Aabb* __restrict__ tri_boxes, const Aabb* __restrict__ node_boxes,
                                  const int* __restrict__ left, const int* __restrict__ right, float4* __restrict__ nodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int lc = left[i], rc = right[i];
  Aabb c[2];
  c[0] = (lc < 0) ? tri_boxes[vals[~lc]] : node_boxes[lc];
  c[1] = (rc < 0) ? tri_boxes[vals[~rc]] : node_boxes[rc];
  const int gamma = (lc < 0) ? ~lc : lc;
  float org[3]; unsigned int eb[3]; float step[3];
  for (int k = 0; k < 3; k++) {
    org[k] = fminf(c[0].mn[k], c[1].mn[k]);
    const float hi = fmaxf(c[0].mx[k], c[1].mx[k]);
    eb[k] = quant_exponent(hi - org[k]);
    step[k] = __uint_as_float(eb[k] << 23);
    // the top of the box must be reachable with 255 steps under the float arithmetic of the decode
    while (eb[k] < 254u && !(org[k] + 255.0f * step[k] >= hi)) { eb[k]++; step[k] = __uint_as_float(eb[k] << 23); }
  }
  unsigned int q[12];
  for (int j = 0; j < 2; j++)
    for (int k = 0; k < 3; k++) {
      int lo = (int)floorf((c[j].mn[k] - org[k]) / step[k]);
      lo = min(max(lo, 0), 255);
      while (lo > 0 && !(org[k] + (float)lo * step[k] <= c[j].mn[k])) lo--;
      int hi = (int)ceilf((c[j].mx[k] - org[k]) / step[k]);
      hi = min(max(hi, 0), 255);
      while (hi < 255 && !(org[k] + (float)hi * step[k] >= c[j].mx[k])) hi++;
      q[6 * j + k] = (unsigned int)lo; q[6 * j + 3 + k] = (unsigned int)hi;
    }
  const unsigned int flags = (lc < 0 ? 1u : 0u) | (rc < 0 ? 2u : 0u);
#if LBVH_NODE_BYTES == 64
  nodes[4 * (size_t)i + 0] = make_float4(c[0].mn[0], c[0].mn[1], c[0].mn[2], c[0].mx[0]);
  nodes[4 * (size_t)i + 1] = make_float4(c[0].mx[1], c[0].mx[2], c[1].mn[0], c[1].mn[1]);
  nodes[4 * (size_t)i + 2] = make_float4(c[1].mn[2], c[1].mx[0], c[1].mx[1], c[1].mx[2]);
  nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(gamma), __uint_as_float(flags), 0.0f, 0.0f);
  return;
#endif
  const unsigned int w0 = eb[0] | (eb[1] << 8) | (eb[2] << 16) | (flags << 24);
  const unsigned int w1 = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
  const unsigned int w2 = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
  const unsigned int w3 = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
  nodes[2 * (size_t)i + 0] = make_float4(org[0], org[1], org[2], __uint_as_float(w0));
  nodes[2 * (size_t)i + 1] = make_float4(__uint_as_float(w1), __uint_as_float(w2), __uint_as_float(w3), __int_as_float(gamma));
}

__global__ void pack_tris_kernel(int n, const int* __restrict__ vals, const float* __restrict__ verts, const int* __restrict__ tris,
                                 float4* __restrict__ out) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  int id = vals[slot];
  int ia = tris[3 * (size_t)id], ib = tris[3 * (size_t)id + 1], ic = tris[3 * (size_t)id + 2];
  float ax = verts[3 * (size_t)ia], ay = verts[3 * (size_t)ia + 1], az = verts[3 * (size_t)ia + 2];
  float bx = verts[3 * (size_t)ib], by = verts[3 * (size_t)ib + 1], bz = verts[3 * (size_t)ib + 2];
  float cx = verts[3 * (size_t)ic], cy = verts[3 * (size_t)ic + 1], cz = verts[3 * (size_t)ic + 2];
  out[3 * (size_t)slot + 0] = make_float4(ax, ay, az, __int_as_float(id));
  out[3 * (size_t)slot + 1] = make_float4(bx - ax, by - ay, bz - az, 0.0f);      // edge1 := B - A
  out[3 * (size_t)slot + 2] = make_float4(cx - ax, cy - ay, cz - az, 0.0f);      // edge2 := C - A
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); gcore_dev::lbvh_free(out); return e_; } } while (0)

}  // namespace

namespace gcore_dev {

void lbvh_free(LbvhBuffers* b) {
  if (!b) return;
  if (b->nodes) cudaFree(b->nodes);
  if (b->tris) cudaFree(b->tris);
  if (b->tri_normals) cudaFree(b->tri_normals);
  if (b->tri_mat) cudaFree(b->tri_mat);
  *b = LbvhBuffers();
}

cudaError_t lbvh_build(const float* d_verts, int nverts, const int* d_tris, int ntris, const float* d_normals9, const int* d_tri_mat,
                       LbvhBuffers* out, cudaStream_t s) {
  (void)nverts;
  lbvh_free(out);
  if (ntris <= 0) return cudaSuccess;
  const int n = ntris;
  Aabb *tri_boxes = nullptr, *node_boxes = nullptr;
  float* scene_bounds = nullptr;
  unsigned long long *keys = nullptr, *keys_sorted = nullptr;
  int *vals = nullptr, *vals_sorted = nullptr, *left = nullptr, *right = nullptr, *par_i = nullptr, *par_l = nullptr, *flags = nullptr;
  void* cub_tmp = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(tri_boxes); cudaFree(node_boxes); cudaFree(scene_bounds); cudaFree(keys); cudaFree(keys_sorted); cudaFree(vals);
    cudaFree(vals_sorted); cudaFree(left); cudaFree(right); cudaFree(par_i); cudaFree(par_l); cudaFree(flags); cudaFree(cub_tmp);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  };
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  CK(cudaMalloc(&tri_boxes, sizeof(Aabb) * (size_t)n));
  CK(cudaMalloc(&scene_bounds, 6 * sizeof(float)));
  CK(cudaMalloc(&keys, 8 * (size_t)n)); CK(cudaMalloc(&keys_sorted, 8 * (size_t)n));
  CK(cudaMalloc(&vals, 4 * (size_t)n)); CK(cudaMalloc(&vals_sorted, 4 * (size_t)n));
  CK(cudaMalloc(&out->tris, 48 * (size_t)n));
  out->ntris = n;
  out->bytes = 48 * (size_t)n;
  if (d_normals9) {
    CK(cudaMalloc(&out->tri_normals, 36 * (size_t)n));
    CK(cudaMemcpyAsync(out->tri_normals, d_normals9, 36 * (size_t)n, cudaMemcpyDeviceToDevice, s));
    out->bytes += 36 * (size_t)n;
  }
  if (d_tri_mat) {
    CK(cudaMalloc(&out->tri_mat, 4 * (size_t)n));
    CK(cudaMemcpyAsync(out->tri_mat, d_tri_mat, 4 * (size_t)n, cudaMemcpyDeviceToDevice, s));
    out->bytes += 4 * (size_t)n;
  }
  const int B = 256, G = (n + B - 1) / B;
  CK(cudaEventRecord(ev0, s));
  init_bounds_kernel<<<1, 32, 0, s>>>(scene_bounds);
  tri_bounds_kernel<<<G, B, 0, s>>>(d_verts, d_tris, n, tri_boxes, scene_bounds);
  morton_kernel<<<G, B, 0, s>>>(tri_boxes, n, scene_bounds, keys, vals);
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, s));
  CK(cudaMalloc(&cub_tmp, tmp_bytes));
  CK(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, s));
  pack_tris_kernel<<<G, B, 0, s>>>(n, vals_sorted, d_verts, d_tris, out->tris);
  if (n >= 2) {
    CK(cudaMalloc(&node_boxes, sizeof(Aabb) * (size_t)(n - 1)));
    CK(cudaMalloc(&left, 4 * (size_t)(n - 1))); CK(cudaMalloc(&right, 4 * (size_t)(n - 1)));
    CK(cudaMalloc(&par_i, 4 * (size_t)(n - 1))); CK(cudaMalloc(&par_l, 4 * (size_t)n));
    CK(cudaMalloc(&flags, 4 * (size_t)(n - 1)));
    CK(cudaMemsetAsync(flags, 0, 4 * (size_t)(n - 1), s));
    CK(cudaMalloc(&out->nodes, LBVH_NODE_BYTES * (size_t)(n - 1)));
    out->nnodes = n - 1; out->root = 0; out->bytes += LBVH_NODE_BYTES * (size_t)(n - 1);
    const int G1 = (n - 1 + B - 1) / B;
    karras_kernel<<<G1, B, 0, s>>>(keys_sorted, n, left, right, par_i, par_l);
    refit_kernel<<<G, B, 0, s>>>(n, vals_sorted, tri_boxes, left, right, par_i, par_l, node_boxes, flags);
    pack_nodes_kernel<<<G1, B, 0, s>>>(n, vals_sorted, tri_boxes, node_boxes, left, right, out->nodes);
  } else {
    out->nnodes = 0; out->root = -1;
  }
  CK(cudaEventRecord(ev1, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&out->build_ms, ev0, ev1);
  cleanup();
  return cudaSuccess;
}

}  // namespace gcore_dev
