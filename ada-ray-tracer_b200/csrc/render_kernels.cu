// render_kernels.cu -- the render hot path on the GPU (sm_100a).
//
// Built with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see art_math.cuh): every result is bit-identical
// to the reference arithmetic restated by the oracle.
//
// Kernels
//   render_kernel<RT,MESHK>  K1+K2+K4+K5+K6a.  One thread per pixel slot, persistent over all samples of the call.  The
//                       per-lane state machine issues exactly ONE scene query per loop iteration -- a camera/bounce ray
//                       (phase BOUNCE) or the shadow ray of the previous hit (phase SHADOW) -- so all lanes of a warp
//                       meet at the single inlined copy of find_closest_hit whatever their phase, bounce or sample
//                       index (path regeneration: a lane starts its next sample as soon as its path ends).  Fused: ray
//                       generation, scene query, light sampling + shadow test, BSDF eval/sample, MIS, recursion unwind,
//                       per-pass accumulation.  Replaces ray_tracer-integrators.adb:25-301, ray_tracer.adb:61-132.
//   debug_kernel        RT_DEBUG / RT_WHITTED palette render, ray_tracer.adb:208-238.
//   primary_kernel      primary hit records (config C2).
//   closest_kernel      Scene.Find_Closest_Hit for a batch of rays.
//   bvh_batch_kernel    triangle-only LBVH query (legacy gcore_closest_hit seam).
//   resolve_kernel      K6b/K7: slot-major accumulator(s) -> row-major HDR float4 and LDR u32 (ray_tracer.adb:281-291).
#include "gcore_internal.h"
#include "art_shade.cuh"
#include "render_wave.cuh"

using namespace art;

namespace {

#ifndef RENDER_BLOCK
#define RENDER_BLOCK 128
#endif
#ifndef RENDER_MINB
#define RENDER_MINB 6        // resident blocks per SM the register allocation is bounded for (6 x 128 thr -> <= 85 regs); measured best of 4..10
#endif
#define MAX_DEPTH_CAP 16

struct PixelId { int x, y, active; };

// slot -> pixel.  A tile is 32x32 pixels = 1024 slots = 32 warps; warp w of a tile covers the 8x4 pixel block
// (w % 4, w / 4); lanes are row-major inside the block.  Slots of a warp are contiguous -> coalesced float4 I/O.
__device__ __forceinline__ PixelId slot_to_pixel(const gcore_dev::RenderArgs& a, int slot, int rank) {
  int tile_local = slot >> 10, q = slot & 1023;
  int tile = tile_local * a.world + rank;
  int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
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
  int tile = ty * a.tiles_x + tx;
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

template <int RT, int MESHK>
__global__ void __launch_bounds__(RENDER_BLOCK, RENDER_MINB) render_kernel(const __grid_constant__ gcore_dev::RenderArgs a) {
  const int slot = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  unsigned int nrays = 0, nmesh = 0, nbounce = 0, nhoisted = 0;
  const PixelId px = slot_to_pixel(a, slot < a.nslots ? slot : 0, a.rank);
  const bool was_live = (slot < a.nslots) && px.active;
  {
    float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (was_live) acc4 = a.accum[slot];
    f3 acc = mk3(acc4.x, acc4.y, acc4.z);
    const f3 bg = mk3(a.bg[0], a.bg[1], a.bg[2]);
    const int aa = a.anti_aliasing;
    const int total = a.passes * (aa ? 4 : 1);
    const int max_depth = a.max_depth;
    const uint32_t pixel = (uint32_t)(px.y * a.width + px.x);

    f3 E[MAX_DEPTH_CAP], Wt[MAX_DEPTH_CAP];     // per-bounce explicit term and continuation weight (local memory)
    f3 Lprim[4];                                // radiance of the sub-pixel camera rays known to end at depth 0
    unsigned int term_mask = 0;                 // bit i: camera ray of sub-pixel i misses / ends on a light
    f3 color = bg;
    f3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);   // the path ray of the next scene query
    f3 Lterm = mk3(0.f, 0.f, 0.f);
    float prev_pdf = 1.0f;
    bool  prev_spec = true;
    int   s = 0, depth = 0;
    bool  live = was_live, has_ray = false, fold_pending = false;

    // Warp-uniform loop, one bounce of every live lane per iteration, all lanes in the same stage at the same time:
    //   (regenerate) -> A path query -> B light sample -> C shadow query -> D NEE + BSDF sample.
    // A lane whose path ended idles through B-D of that iteration and regenerates at the top of the next one; a lane
    // out of samples idles (live = false) instead of leaving, and there is no break/continue in divergent code, so
    // every branch re-converges inside the body.
    while (__any_sync(0xffffffffu, live)) {
      if (live && !has_ray) {
        if (fold_pending) {                                                   // unwind the recursion: integrators.adb:124,191,299
          f3 L = Lterm;
#pragma unroll 1
          for (int k = depth - 1; k >= 0; k--) {
            if (RT == GCORE_PT_STUPID) L = Wt[k] * L;
            else L = E[k] + Wt[k] * L;
          }
          const int sub = aa ? (s & 3) : 0;
          if (depth == 0) { term_mask |= 1u << sub; Lprim[sub] = L; }        // this camera ray ends at depth 0 in every pass
          if (aa) {
            if (sub == 0) color = bg;                                         // :42
            color = color + L;                                                // :47
            if (sub == 3) acc = color + acc;                                  // :51
          } else {
            acc = L + acc;                                                    // :60-63
          }
          s++;
          fold_pending = false;
        }
        // Samples whose camera ray is already known to end at depth 0: the 4 sub-pixel directions are fixed
        // (ray_tracer.adb:72-97), so that query is loop-invariant over the passes of this call; its radiance is added
        // in the reference's order without re-tracing it.
#pragma unroll 1
        while (s < total && ((term_mask >> (aa ? (s & 3) : 0)) & 1u)) {
          const int sub = aa ? (s & 3) : 0;
          const f3 L = Lprim[sub];
          if (aa) {
            if (sub == 0) color = bg;
            color = color + L;
            if (sub == 3) acc = color + acc;
          } else {
            acc = L + acc;
          }
          s++; nhoisted++;
        }
        if (s == total) live = false;
        else {
          o = ld3(g_scn.cam_pos);
          d = eye_ray_dir(a, px.x, px.y, aa ? (s & 3) : -1);
          depth = 0; prev_pdf = 1.0f; prev_spec = true;                       // StartSample, materials.ads:25
          has_ray = true;
        }
      }
      __syncwarp();

      // ---- A: path query (camera or bounce ray), all live lanes
      bool surf = false;
      DHit h;
      if (live) {
        find_closest_hit<MESHK>(o, d, true, h, nmesh);
        nrays++;
        if (!h.is_hit) {                                                      // integrators.adb:98,151,218
          Lterm = mk3(0.f, 0.f, 0.f); has_ray = false; fold_pending = true;
        } else if (g_scn.mats[h.mat].type == MAT_LIGHT) {
          has_ray = false; fold_pending = true;
          if (RT == GCORE_PT_SHADOW) {
            Lterm = mk3(0.f, 0.f, 0.f);                                       // :155-157
          } else if (dot(neg(d), h.normal) < 0.0f) {
            Lterm = mk3(0.f, 0.f, 0.f);                                       // :103, :224
          } else {
            f3 emit = (g_scn.mats[h.mat].light < 0) ? mk3(0.f, 0.f, 0.f) : ld3(g_scn.l_intensity);   // materials.adb:147-156
            if (RT == GCORE_PT_STUPID) Lterm = emit;                          // :106
            else {                                                            // :230-243
              float lgtPdf = light_eval_pdf(o, d, h.t);
              float bsdfPdf = prev_pdf;
              float misWeight = prev_spec ? 1.0f : (bsdfPdf * bsdfPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf));
              Lterm = emit * misWeight;
            }
          }
        } else surf = true;
      }

      // ---- B: light sample + shadow ray (Compute_Shadow, ray_tracer.adb:100-132), lanes standing on a surface
      f3 hpos = o, sdir = d;
      float ls_pdf = 1.0f, sh_thi = 0.0f, sh_tlo = 0.0f, r1 = 0.0f, r2 = 0.0f;
      f3 so = o;
      if (surf) {
        nbounce++;
        const BounceRand br = bounce_rand(a.seed, pixel, (uint32_t)(a.first_sample + s), (uint32_t)depth);
        hpos = mk3(o.x + h.t * d.x, o.y + h.t * d.y, o.z + h.t * d.z);
        r1 = br.u2; r2 = br.u3;
        if (RT != GCORE_PT_STUPID) {                                          // :161-178 / :251-287
          LightSample ls = light_sample(br.u0, br.u1, hpos);
          ls_pdf = ls.pdf;
          float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
          epsilon = max2(epsilon, 1.0e-30f);
          sdir = normalize(ls.pos - hpos);                                    // == normalize(lsam.pos - hpos) of :165/:268
          so = hpos + sdir * epsilon;
          float maxDist  = length(hpos - ls.pos);
          float epsilon2 = max2(maxDist * 0.000001f, 1.0e-30f);
          sh_thi = maxDist - epsilon2;
          sh_tlo = 10.0f * epsilon;
        }
      }

      // ---- C: shadow query (closest hit, only is_hit / t are read: ray_tracer.adb:117-122)
      bool shadowed = false;
      if (RT != GCORE_PT_STUPID) {
        if (surf) {
          DHit hs;
          find_closest_hit<MESHK>(so, sdir, false, hs, nmesh);
          nrays++;
          shadowed = hs.is_hit && (hs.t < sh_thi && hs.t > sh_tlo);
        }
      }

      // ---- D: explicit light term + BSDF sample + next ray
      if (surf) {
        const DMaterial& m = g_scn.mats[h.mat];
        const f3 hn = h.normal;
        const f3 refl = reflect(d, hn);           // shared by Phong eval/sample, mirror, glass (materials.adb:251,313,371,394)
        f3 explicitColor = mk3(0.f, 0.f, 0.f);
        if (RT != GCORE_PT_STUPID) {
          if (!shadowed) {
            f3 bxdfVal; float bsdfPdf;
            mat_eval(m, sdir, hn, refl, bxdfVal, bsdfPdf);
            float cosTheta1 = max2(dot(sdir, hn), 0.0f);
            f3 intensity = ld3(g_scn.l_intensity);
            if (RT == GCORE_PT_SHADOW) {
              explicitColor = intensity * (cosTheta1 * bxdfVal) * (1.0f / max2(ls_pdf, 1.0e-20f));
            } else {
              float lgtPdf = ls_pdf;
              float misWeight = lgtPdf * lgtPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf);
              explicitColor = intensity * (1.0f / max2(lgtPdf, 1.0e-20f)) * (cosTheta1 * bxdfVal) * misWeight;
            }
          }
        }
        MatSampleD ms = mat_sample(m, r1, r2, d, hn, refl);                   // pick up next ray, :116,183,291
        f3 bxdfVal = ms.color * (1.0f / max2(ms.pdf, 1.0e-20f));
        float cosTheta = dot(ms.direction, hn);
        E[depth]  = explicitColor;
        Wt[depth] = fabsf(cosTheta) * bxdfVal;
        depth++;
        f3 off = (sign_ada(cosTheta) * hn) * 1.0e-5f;                         // :122,189,297
        o = hpos + off;
        d = ms.direction;
        prev_pdf = ms.pdf; prev_spec = ms.pureSpecular;
        if (depth == max_depth) {                                             // recursion_level = 0 -> (0,0,0)
          Lterm = mk3(0.f, 0.f, 0.f); has_ray = false; fold_pending = true;
        }
      }
    }
    if (was_live) a.accum[slot] = make_float4(acc.x, acc.y, acc.z, acc4.w);
  }
  // statistics: one atomic per warp and counter
  unsigned int total_rays = __reduce_add_sync(0xffffffffu, nrays);
  unsigned int total_mesh = __reduce_add_sync(0xffffffffu, nmesh);
  unsigned int total_bnc  = __reduce_add_sync(0xffffffffu, nbounce);
  unsigned int total_hst  = __reduce_add_sync(0xffffffffu, nhoisted);
  if ((threadIdx.x & 31) == 0 && (total_rays | total_hst)) {
    atomicAdd(&a.counters[0], (unsigned long long)total_rays);
    atomicAdd(&a.counters[1], (unsigned long long)total_mesh);
    atomicAdd(&a.counters[2], (unsigned long long)total_bnc);
    atomicAdd(&a.counters[3], (unsigned long long)total_hst);
  }
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

__global__ void __launch_bounds__(RENDER_BLOCK) bvh_batch_kernel(int n, const gcore_ray* __restrict__ rays, float t_near, float t_far,
                                                                 gcore_HitCpp* __restrict__ hits, int* __restrict__ found) {
  const int i = blockIdx.x * RENDER_BLOCK + threadIdx.x;
  if (i >= n) return;
  int prim; float t, u, v;
  const f3 o = ld3(rays[i].origin), d = ld3(rays[i].direction);
  bvh_closest(g_scn.bvh_legacy, o, d, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), t_near, t_far, prim, t, u, v);
  found[i] = (prim >= 0) ? 1 : 0;
  if (prim >= 0) {
    gcore_HitCpp hc;
    hc.primIndex = prim; hc.geomIndex = 0; hc.instIndex = 0; hc.t = t;
    hc.normal[0] = 0.f; hc.normal[1] = 0.f; hc.normal[2] = 0.f;
    hc.texCoord[0] = v; hc.texCoord[1] = u;                 // reference v weighs B, u weighs C (geometry.adb:245-246, 301)
    hits[i] = hc;
  }
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

}  // namespace

namespace gcore_dev {

cudaError_t upload_scene(const art::DeviceScene* host_copy, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(g_scn, host_copy, sizeof(art::DeviceScene), 0, cudaMemcpyHostToDevice, s);
}

static inline int blocks_for(int n, int b) { return (n + b - 1) / b; }

template <int MESHK>
static cudaError_t launch_render_m(int render_type, const RenderArgs& a, cudaStream_t s) {
  dim3 g(blocks_for(a.nslots, RENDER_BLOCK)), b(RENDER_BLOCK);
  switch (render_type) {
    case GCORE_PT_STUPID: render_kernel<GCORE_PT_STUPID, MESHK><<<g, b, 0, s>>>(a); break;
    case GCORE_PT_SHADOW: render_kernel<GCORE_PT_SHADOW, MESHK><<<g, b, 0, s>>>(a); break;
    case GCORE_PT_MIS:    render_kernel<GCORE_PT_MIS, MESHK><<<g, b, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_render(int render_type, int mesh_kind, const RenderArgs& a, cudaStream_t s) {
  if (a.nslots <= 0) return cudaSuccess;
  return (mesh_kind == 2) ? launch_render_m<MESHK_BVH>(render_type, a, s) : launch_render_m<MESHK_CONST>(render_type, a, s);
}

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

cudaError_t launch_bvh_batch(int n, const gcore_ray* d_rays, float t_near, float t_far, gcore_HitCpp* d_hits, int* d_found, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  bvh_batch_kernel<<<blocks_for(n, RENDER_BLOCK), RENDER_BLOCK, 0, s>>>(n, d_rays, t_near, t_far, d_hits, d_found);
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

// ---- block-synchronous wavefront render (render_wave.cuh) ----
template <int RT, int MESHK>
static cudaError_t launch_wave_t(const art::WaveArgs& wa, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(art::render_wave_kernel<RT, MESHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(art::WaveSmem));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  art::render_wave_kernel<RT, MESHK><<<grid, WAVE_THREADS, sizeof(art::WaveSmem), s>>>(wa);
  return cudaGetLastError();
}

template <int MESHK>
static cudaError_t launch_wave_m(int render_type, const art::WaveArgs& wa, int grid, cudaStream_t s) {
  switch (render_type) {
    case GCORE_PT_STUPID: return launch_wave_t<GCORE_PT_STUPID, MESHK>(wa, grid, s);
    case GCORE_PT_SHADOW: return launch_wave_t<GCORE_PT_SHADOW, MESHK>(wa, grid, s);
    case GCORE_PT_MIS:    return launch_wave_t<GCORE_PT_MIS, MESHK>(wa, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

int wave_slots_per_block() { return WAVE_NS; }
int wave_blocks_per_sm() { return WAVE_MINB; }
int wave_max_depth() { return WAVE_MAX_DEPTH; }

cudaError_t launch_render_wave(int render_type, int mesh_kind, const RenderArgs& a, const WaveBuffers& wb, int units, cudaStream_t s) {
  if (a.nslots <= 0 || units <= 0) return cudaSuccess;
  art::WaveArgs wa;
  wa.r = a; wa.colbuf = wb.colbuf; wa.stackA = wb.stackA; wa.stackB = wb.stackB; wa.next_item = wb.next_item; wa.units = units;
  cudaError_t e = cudaMemsetAsync(wb.next_item, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  e = (mesh_kind == 2) ? launch_wave_m<MESHK_BVH>(render_type, wa, wb.grid, s) : launch_wave_m<MESHK_CONST>(render_type, wa, wb.grid, s);
  if (e != cudaSuccess) return e;
  art::fold_kernel<<<blocks_for(a.nslots, 256), 256, 0, s>>>(a, wb.colbuf, units);
  return cudaGetLastError();
}

int render_kernel_regs(int render_type) {
  cudaFuncAttributes at;
  cudaError_t e;
  switch (render_type) {
    case GCORE_PT_STUPID: e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_STUPID, MESHK_CONST>); break;
    case GCORE_PT_SHADOW: e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_SHADOW, MESHK_CONST>); break;
    default:              e = cudaFuncGetAttributes(&at, art::render_wave_kernel<GCORE_PT_MIS, MESHK_CONST>); break;
  }
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return at.numRegs;
}

}  // namespace gcore_dev
