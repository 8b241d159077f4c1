// render_kernels.cu -- the render hot path on the GPU (sm_100a).
//
// Built with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see art_math.cuh): every result is bit-identical
// to the reference arithmetic restated by the oracle.
//
// Kernels
//   render_wave_kernel  (render_wave.cuh) the integrator loop as a block-synchronous wavefront: K1+K2+K4+K5+K6a
//   fold_kernel         (render_wave.cuh) per-pass colours -> accumulator, in pass order
//   debug_kernel        RT_DEBUG / RT_WHITTED palette render, ray_tracer.adb:208-238.
//   primary_kernel      primary hit records (config C2).
//   closest_kernel      Scene.Find_Closest_Hit for a batch of rays.
//   bvh_batch_kernel    triangle-only LBVH query (legacy gcore_closest_hit seam).
//   resolve_kernel      K6b/K7: slot-major accumulator(s) -> row-major HDR float4 and LDR u32 (ray_tracer.adb:281-291).
#include "gcore_internal.h"
#include "art_shade.cuh"
#include "render_wave.cuh"
#include <cstddef>

using namespace art;

namespace {

#ifndef RENDER_BLOCK
#define RENDER_BLOCK 128
#endif

struct PixelId { int x, y, active; };

// slot -> pixel.  A tile is 32x32 pixels = 1024 slots = 32 warps; warp w of a tile covers the 8x4 pixel block
// (w % 4, w / 4); lanes are row-major inside the block.  Slots of a warp are contiguous -> coalesced float4 I/O.
__device__ __forceinline__ PixelId slot_to_pixel(const gcore_dev::RenderArgs& a, int slot, int rank) {
  int tile_local = slot >> 10, q = slot & 1023;
  int tile = tile_local * a.world + rank;
  int tx, ty; gcore_dev::gc_tile_to_xy(tile, a.tiles_x, tx, ty);
  int w = q >> 5, l = q & 31;
  PixelId p;
  p.x = tx * GC_TILE + (w & 3) * 8 + (l & 7);
  p.y = ty * GC_TILE + (w >> 2) * 4 + (l >> 3);
  p.active = (tile < a.ntiles) && (p.x < a.width) && (p.y < a.height);
  return p;
}

// pixel -> (rank, slot) for the resolve
__device__ __forceinline__ void pixel_to_slot(const gcore_dev::RenderArgs& a, int x, int y, int& rank, int& slot) {
  int tx = x / GC_TILE, ty = y / GC_TILE, lx = x % GC_TILE, ly = y % GC_TILE;
  int tile = gcore_dev::gc_xy_to_tile(tx, ty, a.tiles_x);
  rank = tile % a.world;
  int tile_local = tile / a.world;
  int w = (ly >> 2) * 4 + (lx >> 3), l = (ly & 3) * 8 + (lx & 7);
  slot = (tile_local << 10) + (w << 5) + l;
}

// EyeRayDirection / Generate4RayDirections (ray_tracer.adb:61-97) + normalize(cam.matrix * dir) (integrators.adb:46,58)
__device__ __forceinline__ f3 eye_ray_dir(const gcore_dev::RenderArgs& a, int x, int y, int sub) {
  const float third = 1.0f / 3.0f, two_thirds = 2.0f / 3.0f;
  float offx = (sub < 0) ? 0.5f : ((sub & 2) ? two_thirds : third);
  float offy = (sub < 0) ? 0.5f : ((sub & 1) ? two_thirds : third);
  f3 r;
  r.x = (float)x + offx - a.half_w;
  r.y = (float)y + offy - a.half_h;
  r.z = a.zdir;
  r = normalize(r);
  f3 m;
  m.x = g_scn.cam_m[0] * r.x + g_scn.cam_m[1] * r.y + g_scn.cam_m[2] * r.z + g_scn.cam_m[3];   // vector_math.adb:137-144
  m.y = g_scn.cam_m[4] * r.x + g_scn.cam_m[5] * r.y + g_scn.cam_m[6] * r.z + g_scn.cam_m[7];
  m.z = g_scn.cam_m[8] * r.x + g_scn.cam_m[9] * r.y + g_scn.cam_m[10] * r.z + g_scn.cam_m[11];
  return normalize(m);
}

__device__ __forceinline__ void store_hit(gcore_hit* o, const DHit& h) {
  o->is_hit = h.is_hit; o->prim_type = h.prim_type; o->prim_index = h.prim_index; o->matId = h.matId; o->mat = h.mat;
  o->t = h.t; o->normal[0] = h.normal.x; o->normal[1] = h.normal.y; o->normal[2] = h.normal.z; o->tx = h.tx; o->ty = h.ty;
}

template <int MESHK>
__global__ void __launch_bounds__(RENDER_BLOCK) debug_kernel(const __grid_constant__ gcore_dev::RenderArgs a) {
  const int slot = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  unsigned int nrays = 0;
  const PixelId px = slot_to_pixel(a, slot < a.nslots ? slot : 0, a.rank);
  if (slot < a.nslots && px.active) {
    DHit h;
    unsigned int nmesh = 0;
    find_closest_hit<MESHK>(ld3(g_scn.cam_pos), eye_ray_dir(a, px.x, px.y, -1), true, h, nmesh);
    nrays = 1;
    float3 c = make_float3(0.f, 0.f, 0.f);
    if (h.is_hit) {                                                           // palette, ray_tracer.adb:210-211
      switch (h.matId % 8) {
        case 0: c = make_float3(0.5f, 0.0f, 0.0f); break;
        case 1: c = make_float3(0.0f, 0.5f, 0.0f); break;
        case 2: c = make_float3(0.0f, 0.0f, 0.5f); break;
        case 3: c = make_float3(0.5f, 0.5f, 0.5f); break;
        case 4: c = make_float3(0.5f, 0.5f, 0.0f); break;
        case 5: c = make_float3(0.5f, 0.0f, 0.5f); break;
        case 6: c = make_float3(0.0f, 0.5f, 0.5f); break;
        default: c = make_float3(0.75f, 0.75f, 0.75f); break;
      }
    }
    a.accum[slot] = make_float4(c.x, c.y, c.z, 0.f);
  }
  unsigned int total_rays = __reduce_add_sync(0xffffffffu, nrays);
  if ((threadIdx.x & 31) == 0 && total_rays) atomicAdd(&a.counters[0], (unsigned long long)total_rays);
}

template <int MESHK>
__global__ void __launch_bounds__(RENDER_BLOCK) primary_kernel(const __grid_constant__ gcore_dev::RenderArgs a, gcore_hit* __restrict__ hits) {
  const int slot = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  if (slot >= a.nslots) return;
  const PixelId px = slot_to_pixel(a, slot, a.rank);
  if (!px.active) return;
  DHit h;
  unsigned int nmesh = 0;
  find_closest_hit<MESHK>(ld3(g_scn.cam_pos), eye_ray_dir(a, px.x, px.y, -1), true, h, nmesh);
  store_hit(&hits[(size_t)px.y * a.width + px.x], h);
}

template <int MESHK>
__global__ void __launch_bounds__(RENDER_BLOCK) closest_kernel(int n, const gcore_ray* __restrict__ rays, gcore_hit* __restrict__ hits) {
  const int i = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  if (i >= n) return;
  DHit h;
  unsigned int nmesh = 0;
  find_closest_hit<MESHK>(ld3(rays[i].origin), ld3(rays[i].direction), true, h, nmesh);
  store_hit(&hits[i], h);
}

// STATS: also count the nodes visited / triangle slots tested (counters[0], [1]) and the hits ([2]); which: 0 = the scene
// committed through the legacy seam, 1 = the shaded mesh of the renderer.  TWO_SIDED: Embree's two-sided triangle test instead of
// the reference's one-sided one.  ANY_HIT: occlusion query (found[] only, hits may be null).
template <bool STATS, bool TWO_SIDED, bool ANY_HIT>
__global__ void __launch_bounds__(RENDER_BLOCK) bvh_batch_kernel(int n, const gcore_ray* __restrict__ rays, float t_near, float t_far,
                                                                 gcore_HitCpp* __restrict__ hits, int* __restrict__ found, int which,
                                                                 unsigned long long* __restrict__ counters) {
  const int i = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  const BvhView& bv = which ? g_scn.bvh : g_scn.bvh_legacy;
  int prim = -1, slot = -1; float t = 0.f, u = 0.f, v = 0.f;
  BvhStats st; st.nodes = 0; st.tris = 0;
  f3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);
  if (i < n) {
    o = ld3(rays[i].origin); d = ld3(rays[i].direction);
    bvh_closest_slot_t<STATS, TWO_SIDED, ANY_HIT>(bv, o, d, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), t_near, t_far, prim, t, u, v, slot, st);
  }
  if (STATS) {
    unsigned int a = __reduce_add_sync(0xffffffffu, st.nodes), b = __reduce_add_sync(0xffffffffu, st.tris);
    unsigned int c = __reduce_add_sync(0xffffffffu, (unsigned int)(prim >= 0));
    if ((threadIdx.x & 31) == 0) { atomicAdd(&counters[0], (unsigned long long)a); atomicAdd(&counters[1], (unsigned long long)b); atomicAdd(&counters[2], (unsigned long long)c); }
  }
  if (i >= n) return;
  found[i] = (prim >= 0) ? 1 : 0;
  if (!ANY_HIT && prim >= 0) {
    gcore_HitCpp hc;
    hc.primIndex = prim; hc.geomIndex = 0; hc.instIndex = 0; hc.t = t;
    const float4 e1 = __ldg(&bv.tris[3 * slot + 1]), e2 = __ldg(&bv.tris[3 * slot + 2]);
    const f3 ng = cross(mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z));       // Ng = (B - A) x (C - A), unnormalised like Embree's
    hc.normal[0] = ng.x; hc.normal[1] = ng.y; hc.normal[2] = ng.z;
    hc.texCoord[0] = v; hc.texCoord[1] = u;                 // reference v weighs B, u weighs C (geometry.adb:245-246, 301)
    hits[i] = hc;
  }
}

// the one-ray seam (gcore_closest_hit): a combined batch of the calls that are pending across the host's threads.  Requests and
// answers live in mapped pinned host memory -- the kernel reads the rays and writes the hit records over PCIe itself, so a batch
// costs one launch and one stream synchronisation, no copies.
static_assert(sizeof(gcore_dev::SeamIn) == 32 && sizeof(gcore_dev::SeamOut) == 40, "the seam kernels move requests / answers as 8 / 10 words");

// one request (8 words: origin, direction, t_near, t_far) -> one answer (10 words: HitCpp, found)
template <bool TWO_SIDED>
__device__ __forceinline__ void seam_answer(const float* in8, unsigned int* out10) {
  const BvhView& bv = g_scn.bvh_legacy;
  const f3 o = mk3(in8[0], in8[1], in8[2]), d = mk3(in8[3], in8[4], in8[5]);
  int prim = -1, slot = -1; float t = 0.f, u = 0.f, v = 0.f;
  BvhStats st; st.nodes = 0; st.tris = 0;
  bvh_closest_slot_t<false, TWO_SIDED, false>(bv, o, d, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), in8[6], in8[7], prim, t, u, v, slot, st);
  f3 ng = mk3(0.f, 0.f, 0.f);
  if (prim >= 0) {
    const float4 e1 = __ldg(&bv.tris[3 * slot + 1]), e2 = __ldg(&bv.tris[3 * slot + 2]);
    ng = cross(mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z));                 // Ng = (B - A) x (C - A), unnormalised like Embree's
  }
  out10[0] = (unsigned)prim; out10[1] = 0u; out10[2] = 0u;                     // primIndex, geomIndex, instIndex (mapped on the host)
  out10[3] = __float_as_uint(t);
  out10[4] = __float_as_uint(ng.x); out10[5] = __float_as_uint(ng.y); out10[6] = __float_as_uint(ng.z);
  out10[7] = __float_as_uint(v); out10[8] = __float_as_uint(u);               // reference v weighs B, u weighs C (geometry.adb:245-246, 301)
  out10[9] = (prim >= 0) ? 1u : 0u;
}

// the one-ray seam (gcore_closest_hit) without the resident kernel: one block of one warp per pending request, so that the handful of
// latency-bound walks spread over the SMs.  The warp reads its request and writes its answer as ONE access each.
template <bool TWO_SIDED>
__global__ void __launch_bounds__(32) bvh_seam_kernel(int n, gcore_dev::SeamBatch* __restrict__ b) {
  __shared__ float s_in[8];
  __shared__ unsigned int s_out[10];
  const int i = blockIdx.x, lane = threadIdx.x;
  if (i >= n) return;
  if (lane < 8) s_in[lane] = ((const volatile float*)&b->in[i])[lane];
  __syncwarp();
  if (lane == 0) seam_answer<TWO_SIDED>(s_in, s_out);
  __syncwarp();
  if (lane < 10) ((volatile unsigned int*)&b->out[i])[lane] = s_out[lane];
}

// One ray, one WARP: the walk of a single ray is a chain of 40-100 dependent fetches when one lane does it, and a batch of the seam is as
// slow as its slowest ray.  Here the warp keeps a frontier of inner nodes in shared memory and every round its lanes take the last (up to)
// 32 entries: one fetch latency per round for 32 boxes pairs, leaves are tested by the lane that finds them, the lanes' best hits are
// merged by (t, then lowest triangle id) -- the closest hit and the tie rule of bvh_closest_slot_t do not depend on the order of the
// visits, and the culling bound is the warp's current minimum (closed, so that equal-t candidates still reach the tie rule).  Near-first
// order is lost inside a round, so more nodes are touched; what shrinks is the number of DEPENDENT rounds (about the depth of the tree).
// A frontier that would not fit (SEAM_FRONTIER entries) sends the ray to the one-lane walk instead.
#define SEAM_FRONTIER 768
#ifndef SEAM_WARP_WALK
#define SEAM_WARP_WALK 1
#endif
template <bool TWO_SIDED>
__device__ __forceinline__ void seam_answer_warp(const float* in8, unsigned int* out10, int* frontier, int lane) {
  const BvhView& bv = g_scn.bvh_legacy;
  const unsigned int FULL = 0xffffffffu;
  const f3 o = mk3(in8[0], in8[1], in8[2]), d = mk3(in8[3], in8[4], in8[5]);
  const float t_near = in8[6], t_far = in8[7];
  if (bv.ntris <= 0 || bv.root < 0) {                                          // empty or one-triangle scene: nothing to share out
    if (lane == 0) seam_answer<TWO_SIDED>(in8, out10);
    __syncwarp();
    return;
  }
  const f3 inv = bvh_slab_inv_of(mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z));
  float best_t = t_far, best_u = 0.f, best_v = 0.f; int best_prim = -1, best_slot = -1;      // this lane's best
  float bound = t_far;                                                         // the warp's best so far
  int count = 1;
  bool overflow = false;
  if (lane == 0) frontier[0] = bv.root;
  __syncwarp();
  while (count > 0) {
    const int take = count < 32 ? count : 32;
    const int node = (lane < take) ? frontier[count - 1 - lane] : -1;
    count -= take;
    __syncwarp();
    int push0 = -1, push1 = -1;
    if (node >= 0) {
      int c0, c1; float e0, e1; bool h0, h1;
      bvh_node_children(bv, node, o, inv, t_near, bound, c0, c1, h0, h1, e0, e1);
      if (h0) { if (c0 < 0) bvh_leaf_test<TWO_SIDED>(bv, ~c0, 1, o, d, t_near, best_t, best_prim, best_u, best_v, best_slot); else push0 = c0; }
      if (h1) { if (c1 < 0) bvh_leaf_test<TWO_SIDED>(bv, ~c1, 1, o, d, t_near, best_t, best_prim, best_u, best_v, best_slot); else push1 = c1; }
      if (push0 >= 0 && push1 >= 0 && e0 < e1) { const int t = push0; push0 = push1; push1 = t; }   // the nearer child is filed last: taken first
    }
    float m = best_t;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, off));
    bound = m;
    const unsigned int m0 = __ballot_sync(FULL, push0 >= 0), m1 = __ballot_sync(FULL, push1 >= 0);
    const int n0 = __popc(m0), n1 = __popc(m1);
    if (count + n0 + n1 > SEAM_FRONTIER) { overflow = true; break; }
    const unsigned int below = (1u << lane) - 1u;
    if (push0 >= 0) frontier[count + __popc(m0 & below)] = push0;
    if (push1 >= 0) frontier[count + n0 + __popc(m1 & below)] = push1;
    count += n0 + n1;
    __syncwarp();
  }
  if (overflow) {                                                              // (uniform) never seen on the test scenes; correctness does not depend on it
    if (lane == 0) seam_answer<TWO_SIDED>(in8, out10);
    __syncwarp();
    return;
  }
  // the winner: smallest t, ties to the lowest original triangle id (the rule inside bvh_leaf_test)
  float wt = best_t; int wp = (best_prim >= 0) ? best_prim : 0x7fffffff; int wl = lane;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ot = __shfl_xor_sync(FULL, wt, off); const int op = __shfl_xor_sync(FULL, wp, off); const int ol = __shfl_xor_sync(FULL, wl, off);
    const bool theirs = (op != 0x7fffffff) && (wp == 0x7fffffff || ot < wt || (ot == wt && op < wp));
    if (theirs) { wt = ot; wp = op; wl = ol; }
  }
  if (wp == 0x7fffffff) {
    if (lane == 0) { for (int k = 0; k < 9; k++) out10[k] = 0u; out10[0] = (unsigned)-1; out10[9] = 0u; }
  } else if (lane == wl) {
    const float4 e1 = __ldg(&bv.tris[3 * best_slot + 1]), e2 = __ldg(&bv.tris[3 * best_slot + 2]);
    const f3 ng = cross(mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z));
    out10[0] = (unsigned)best_prim; out10[1] = 0u; out10[2] = 0u; out10[3] = __float_as_uint(best_t);
    out10[4] = __float_as_uint(ng.x); out10[5] = __float_as_uint(ng.y); out10[6] = __float_as_uint(ng.z);
    out10[7] = __float_as_uint(best_v); out10[8] = __float_as_uint(best_u);
    out10[9] = 1u;
  }
  __syncwarp();
}

// The resident seam: thread 0 polls the mailbox's command word over PCIe; a new sequence number brings a batch of n <= 64 requests in
// batch buffer 0 or 1 (mapped pinned memory, like the mailbox).  The block copies the n x 8 request words into shared memory with
// coalesced reads (volatile: the same host addresses carry a new batch every time), request i is answered by warp i % 32 with all its
// lanes (seam_answer_warp; requests 32.. in a second round) and the n x 10 answer words go back coalesced; then `done` := the sequence number, after a system-wide fence.  Without a command for idle_ns
// (or with `stop` set) the kernel clears `alive` and returns: nothing else waits on it for longer than that.
__device__ __forceinline__ unsigned long long seam_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool TWO_SIDED>
__global__ void __launch_bounds__(1024) bvh_seam_server(gcore_dev::SeamMailbox* mb, gcore_dev::SeamBatch* b0, gcore_dev::SeamBatch* b1,
                                                        unsigned int seen, long long idle_ns) {
  __shared__ unsigned int s_cmd;
  __shared__ float s_in[8 * GC_SEAM_BATCH];
  __shared__ unsigned int s_out[10 * GC_SEAM_BATCH];
  extern __shared__ int s_frontier[];                                         // SEAM_FRONTIER entries per warp
  unsigned long long t_last = seam_now_ns();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) {
      unsigned int c;
      for (;;) {
        c = *(volatile unsigned int*)&mb->cmd;
        if ((c >> 8) != seen) break;
        if (*(volatile unsigned int*)&mb->stop != 0u || (long long)(seam_now_ns() - t_last) > idle_ns) { c = 0xffffffffu; break; }
      }
      s_cmd = c;
    }
    __syncthreads();
    const unsigned int c = s_cmd;
    if (c == 0xffffffffu) break;
    const int n = (int)(c & 127u);
    gcore_dev::SeamBatch* b = (c & 128u) ? b1 : b0;
    if (tid < 8 * n) s_in[tid] = ((const volatile float*)b->in)[tid];
    __syncthreads();
#if SEAM_WARP_WALK
    for (int i = warp; i < n; i += 32) seam_answer_warp<TWO_SIDED>(&s_in[8 * i], &s_out[10 * i], &s_frontier[SEAM_FRONTIER * warp], lane);   // (uniform per warp)
#else
    const int i = warp + 32 * lane;
    if (lane < 2 && i < n) seam_answer<TWO_SIDED>(&s_in[8 * i], &s_out[10 * i]);
#endif
    __syncthreads();
    if (tid < 10 * n) ((volatile unsigned int*)b->out)[tid] = s_out[tid];
    __threadfence_system();                                                   // the answers are in host memory before `done` says so
    __syncthreads();
    seen = c >> 8;
    if (tid == 0) { *(volatile unsigned int*)&mb->done = seen; t_last = seam_now_ns(); }
  }
  if (tid == 0) { __threadfence_system(); *(volatile unsigned int*)&mb->alive = 0u; }
}

// Ada float -> Unsigned_32 conversion rounds to nearest, ties away from zero (ray_tracer.adb:51-53)
__device__ __forceinline__ uint32_t to_u32_round(float v) {
  float f = floorf(v);
  float dlt = v - f;
  if (dlt >= 0.5f) f = f + 1.0f;
  if (!(f > 0.0f)) return 0u;
  return (uint32_t)f;
}

__global__ void __launch_bounds__(256) resolve_kernel(const __grid_constant__ gcore_dev::RenderArgs a, const float4* __restrict__ all,
                                                      int gathered, int spp, float4* __restrict__ hdr4, uint32_t* __restrict__ ldr) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.width || y >= a.height) return;
  int rank, slot;
  pixel_to_slot(a, x, y, rank, slot);
  const float4 v = gathered ? all[(size_t)rank * a.nslots + slot] : all[slot];
  const size_t o = (size_t)y * a.width + x;
  if (hdr4) hdr4[o] = make_float4(v.x, v.y, v.z, (float)spp);
  if (ldr) {
    float normC = 1.0f / (float)spp;                                          // ray_tracer.adb:281
    float r = sqrtf(normC * v.x), g = sqrtf(normC * v.y), b = sqrtf(normC * v.z);   // ** (1.0/g_gamma) with g_gamma = 2.0
    r = min2(r, 1.0f); g = min2(g, 1.0f); b = min2(b, 1.0f);                  // ToneMapping :19-38
    ldr[o] = to_u32_round(r * 255.0f) | (to_u32_round(g * 255.0f) << 8) | (to_u32_round(b * 255.0f) << 16);
  }
}

// checkpoint / resume: row-major HDR frame (gcore_read_frame) -> this rank's slot-major accumulator
__global__ void __launch_bounds__(256) load_accum_kernel(const __grid_constant__ gcore_dev::RenderArgs a, const float4* __restrict__ hdr4) {
  const int slot = blockIdx.x * 256 + threadIdx.x;
  if (slot >= a.nslots) return;
  const PixelId px = slot_to_pixel(a, slot, a.rank);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (px.active) { v = hdr4[(size_t)px.y * a.width + px.x]; v.w = 0.f; }
  a.accum[slot] = v;
}

}  // namespace

namespace gcore_dev {

cudaError_t launch_load_accum(const RenderArgs& a, const float4* d_hdr4, cudaStream_t s) {
  if (a.nslots <= 0) return cudaSuccess;
  load_accum_kernel<<<(a.nslots + 255) / 256, 256, 0, s>>>(a, d_hdr4);
  return cudaGetLastError();
}


cudaError_t upload_scene(const art::DeviceScene* host_copy, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(g_scn, host_copy, sizeof(art::DeviceScene), 0, cudaMemcpyHostToDevice, s);
}

static inline int blocks_for(int n, int b) { return (n + b - 1) / b; }

cudaError_t launch_debug(int mesh_kind, const RenderArgs& a, cudaStream_t s) {
  if (a.nslots <= 0) return cudaSuccess;
  if (mesh_kind == 2) debug_kernel<MESHK_BVH><<<blocks_for(a.nslots, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(a);
  else debug_kernel<MESHK_CONST><<<blocks_for(a.nslots, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_primary_hits(int mesh_kind, const RenderArgs& a, gcore_hit* d_hits, cudaStream_t s) {
  if (a.nslots <= 0) return cudaSuccess;
  if (mesh_kind == 2) primary_kernel<MESHK_BVH><<<blocks_for(a.nslots, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(a, d_hits);
  else primary_kernel<MESHK_CONST><<<blocks_for(a.nslots, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(a, d_hits);
  return cudaGetLastError();
}

cudaError_t launch_closest_batch(int mesh_kind, int n, const gcore_ray* d_rays, gcore_hit* d_hits, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (mesh_kind == 2) closest_kernel<MESHK_BVH><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, d_hits);
  else closest_kernel<MESHK_CONST><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, d_hits);
  return cudaGetLastError();
}

cudaError_t launch_bvh_batch(int n, const gcore_ray* d_rays, float t_near, float t_far, gcore_HitCpp* d_hits, int* d_found, int two_sided, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (two_sided) bvh_batch_kernel<false, true, false><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, d_hits, d_found, 0, nullptr);
  else bvh_batch_kernel<false, false, false><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, d_hits, d_found, 0, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_bvh_seam(int n, SeamBatch* d_batch, int two_sided, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (two_sided) bvh_seam_kernel<true><<<n, 32, 0, s>>>(n, d_batch);
  else bvh_seam_kernel<false><<<n, 32, 0, s>>>(n, d_batch);
  return cudaGetLastError();
}

cudaError_t launch_seam_server(SeamMailbox* d_mb, SeamBatch* d_batch0, SeamBatch* d_batch1, unsigned int seen_seq, long long idle_ns,
                               int two_sided, cudaStream_t s) {
  const int smem = SEAM_FRONTIER * 32 * (int)sizeof(int);                    // the warps' frontiers (96 KB)
  cudaError_t e = two_sided ? cudaFuncSetAttribute(bvh_seam_server<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                            : cudaFuncSetAttribute(bvh_seam_server<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  if (two_sided) bvh_seam_server<true><<<1, 1024, smem, s>>>(d_mb, d_batch0, d_batch1, seen_seq, idle_ns);
  else bvh_seam_server<false><<<1, 1024, smem, s>>>(d_mb, d_batch0, d_batch1, seen_seq, idle_ns);
  return cudaGetLastError();
}

cudaError_t launch_bvh_occluded(int n, const gcore_ray* d_rays, float t_near, float t_far, int* d_occluded, int two_sided, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (two_sided) bvh_batch_kernel<false, true, true><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, nullptr, d_occluded, 0, nullptr);
  else bvh_batch_kernel<false, false, true><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, nullptr, d_occluded, 0, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_bvh_batch_which(int n, const gcore_ray* d_rays, float t_near, float t_far, gcore_HitCpp* d_hits, int* d_found, int which,
                                   unsigned long long* d_counters, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (d_counters) bvh_batch_kernel<true, false, false><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, d_hits, d_found, which, d_counters);
  else bvh_batch_kernel<false, false, false><<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, d_hits, d_found, which, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_resolve(const RenderArgs& a, int spp, float4* d_hdr4, uint32_t* d_ldr, cudaStream_t s) {
  dim3 g(blocks_for(a.width, 32), blocks_for(a.height, 8)), b(256);
  resolve_kernel<<<g, b, 0, s>>>(a, a.accum, 0, spp, d_hdr4, d_ldr);
  return cudaGetLastError();
}

cudaError_t launch_resolve_gathered(const RenderArgs& a, const float4* d_all, int spp, float4* d_hdr4, uint32_t* d_ldr, cudaStream_t s) {
  dim3 g(blocks_for(a.width, 32), blocks_for(a.height, 8)), b(256);
  resolve_kernel<<<g, b, 0, s>>>(a, d_all, 1, spp, d_hdr4, d_ldr);
  return cudaGetLastError();
}

// ---- wavefront render (render_wave.cuh) ----
#ifndef WAVE_NT_CONST
#define WAVE_NT_CONST 1024
#endif

template <int RT>
static cudaError_t launch_wave_const(const art::WaveArgs& wa, int sm_count, cudaStream_t s) {
  constexpr int NT = WAVE_NT_CONST;
  typedef art::WaveSmem<NT> Smem;
  // a per-device attribute (one host process may drive several GPUs): set it with every launch, it costs a microsecond
  cudaError_t e = cudaFuncSetAttribute(art::render_wave_kernel<RT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  const int grid = sm_count * WAVE_BLOCKS_PER_SM(NT);                        // persistent: one wave of resident blocks
  art::render_wave_kernel<RT, NT><<<grid, NT, sizeof(Smem), s>>>(wa);
  return cudaGetLastError();
}

template <int RT>
static cudaError_t launch_wave_bvh(const art::WaveArgs& wa, int sm_count, cudaStream_t s) {
  constexpr int NT = BVH_SHADE_THREADS + BVH_TRAV_THREADS;
  cudaError_t e = cudaFuncSetAttribute(art::render_bvh_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(art::BvhWaveSmem));
  if (e != cudaSuccess) return e;
  const int grid = sm_count * BVH_BLOCKS_PER_SM;
  art::render_bvh_kernel<RT><<<grid, NT, sizeof(art::BvhWaveSmem), s>>>(wa);
  return cudaGetLastError();
}

int wave_total_slots(int sm_count) { return sm_count * WAVE_MAX_SLOTS_PER_SM; }   // upper bound of the path slots resident on a device, either kernel
int wave_max_depth() { return WAVE_MAX_DEPTH; }

cudaError_t launch_render_wave(int render_type, int mesh_kind, const RenderArgs& a, const WaveBuffers& wb, int units, int count_mesh, cudaStream_t s) {
  if (a.nslots <= 0 || units <= 0) return cudaSuccess;
  art::WaveArgs wa;
  wa.r = a; wa.colbuf = wb.colbuf; wa.stackA = wb.stackA; wa.stackB = wb.stackB; wa.next_item = wb.next_item; wa.units = units; wa.count_mesh = count_mesh;
  wa.tailbuf = wb.tailbuf;
  wa.total_items = (unsigned int)units * (unsigned int)a.nslots;
  // the last items of the launch are handed out path by path (4 per item with anti-aliasing on): enough of them to keep every slot busy
  // while the whole items drain
  unsigned int tail_items = a.anti_aliasing ? (unsigned int)wb.tail_items : 0u;
  if (tail_items > wa.total_items) tail_items = wa.total_items;
  wa.tail_first = wa.total_items - tail_items;
  cudaError_t e = cudaMemsetAsync(wb.next_item, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  const bool bvh = (mesh_kind == 2);
  switch (render_type) {
    case GCORE_PT_STUPID: e = bvh ? launch_wave_bvh<GCORE_PT_STUPID>(wa, wb.sm_count, s) : launch_wave_const<GCORE_PT_STUPID>(wa, wb.sm_count, s); break;
    case GCORE_PT_SHADOW: e = bvh ? launch_wave_bvh<GCORE_PT_SHADOW>(wa, wb.sm_count, s) : launch_wave_const<GCORE_PT_SHADOW>(wa, wb.sm_count, s); break;
    case GCORE_PT_MIS:    e = bvh ? launch_wave_bvh<GCORE_PT_MIS>(wa, wb.sm_count, s) : launch_wave_const<GCORE_PT_MIS>(wa, wb.sm_count, s); break;
    default: e = cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  art::fold_kernel<<<blocks_for(a.nslots, 256), 256, 0, s>>>(a, wb.colbuf, wb.tailbuf, wa.tail_first, units);
  return cudaGetLastError();
}

int render_kernel_regs(int render_type) {
  cudaFuncAttributes at;
  cudaError_t e;
  switch (render_type) {
    case GCORE_PT_STUPID: e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_STUPID, WAVE_NT_CONST>); break;
    case GCORE_PT_SHADOW: e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_SHADOW, WAVE_NT_CONST>); break;
    default:              e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_MIS, WAVE_NT_CONST>); break;
  }
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return at.numRegs;
}

}  // namespace gcore_dev
