// gcore_internal.h -- host-side launch interface between gcore_api.cu, render_kernels.cu and lbvh_build.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcore.h"

namespace art { struct DeviceScene; struct BvhView; }

namespace gcore_dev {

#define GC_TILE 32                       // tile edge in pixels; one tile = 1024 accumulator slots
// Tile order: the tiles of row ty are numbered ty * tiles_x + (tx + GC_TILE_SKEW * ty) % tiles_x, and tile number t belongs
// to rank t % world.  The rotation keeps a rank from owning whole tile columns when tiles_x is a multiple of world (a plain
// t % world then gives every rank the same x positions in every row, e.g. a rank with nothing but background).
#define GC_TILE_SKEW 3
__host__ __device__ inline void gc_tile_to_xy(int tile, int tiles_x, int& tx, int& ty) {
  ty = tile / tiles_x;
  int r = tile - ty * tiles_x - (GC_TILE_SKEW * ty) % tiles_x;
  tx = r < 0 ? r + tiles_x : r;
}
__host__ __device__ inline int gc_xy_to_tile(int tx, int ty, int tiles_x) { return ty * tiles_x + (tx + GC_TILE_SKEW * ty) % tiles_x; }

struct RenderArgs {
  int      width, height;
  int      tiles_x, ntiles;              // tile grid of the whole image
  int      rank, world;                  // this device renders tiles with (tile % world) == rank
  int      nslots;                       // local tiles * 1024
  int      anti_aliasing, max_depth, passes;
  uint32_t seed;
  int      first_sample;
  float    bg[3];
  float    zdir;                         // -float(width) / safe_tan(fov/2)   (ray_tracer.adb:67)
  float    half_w, half_h;               // float(width)/2, float(height)/2
  float4*  accum;                        // slot-major accumulator (rgb, unused)
  unsigned long long* counters;          // [0] += scene queries, [1] += queries inside the mesh bbox, [2] += shaded bounces, [3] += LBVH nodes visited, [4] += LBVH triangle slots tested, [5] += LBVH traversals (6 counters)
};

cudaError_t upload_scene(const art::DeviceScene* host_copy, cudaStream_t s);
// RT_DEBUG / RT_WHITTED: palette of the primary hit written into accum (ray_tracer.adb:208-238)
cudaError_t launch_debug(int mesh_kind, const RenderArgs& a, cudaStream_t s);
// primary hits, row-major hits[y*W+x]
cudaError_t launch_primary_hits(int mesh_kind, const RenderArgs& a, gcore_hit* d_hits, cudaStream_t s);
cudaError_t launch_closest_batch(int mesh_kind, int n, const gcore_ray* d_rays, gcore_hit* d_hits, cudaStream_t s);
// pure triangle query on the committed LBVH (legacy gcore_closest_hit semantics)
// two_sided: Embree's rtcIntersect1 semantics (both faces) instead of the reference's one-sided triangle test
cudaError_t launch_bvh_batch(int n, const gcore_ray* d_rays, float t_near, float t_far, gcore_HitCpp* d_hits, int* d_found, int two_sided, cudaStream_t s);
// the combined batch of pending gcore_closest_hit calls: requests / answers in mapped pinned host memory (n <= GC_SEAM_BATCH)
#define GC_SEAM_BATCH 64
// Requests and answers are two dense arrays so that the device moves them over PCIe in a few wide transactions (32 lanes x 4 B) instead
// of eighteen 4-byte ones per ray -- measured: the narrow accesses cost ~4 us per ray of a batch, more than the walks themselves.
struct SeamIn  { float o[3], d[3], t_near, t_far; };            // 8 words
struct SeamOut { gcore_HitCpp hit; int found; };                 // 10 words
struct SeamBatch { SeamIn in[GC_SEAM_BATCH]; SeamOut out[GC_SEAM_BATCH]; };
cudaError_t launch_bvh_seam(int n, SeamBatch* d_batch, int two_sided, cudaStream_t s);
// the resident form of the same kernel: one block that stays on an SM between calls and takes its batches from a mailbox in
// mapped pinned host memory (cmd = seq << 8 | buffer << 7 | n), so that a call costs two PCIe round trips instead of a kernel launch
// and a stream synchronisation.  It leaves by itself after `idle_ns` without a command, or when `stop` is set.
struct SeamMailbox { unsigned int cmd, stop, pad0[14]; unsigned int done, alive, pad1[14]; };   // host-written line, device-written line
cudaError_t launch_seam_server(SeamMailbox* d_mb, SeamBatch* d_batch0, SeamBatch* d_batch1, unsigned int seen_seq, long long idle_ns,
                               int two_sided, cudaStream_t s);
// occlusion query on the committed legacy scene (rtcOccluded1 semantics: any hit inside (t_near, t_far)); d_occluded[i] = 0 / 1
cudaError_t launch_bvh_occluded(int n, const gcore_ray* d_rays, float t_near, float t_far, int* d_occluded, int two_sided, cudaStream_t s);
// the same kernel on either LBVH (which: 0 legacy scene, 1 shaded mesh); d_counters != null: also count nodes visited,
// triangle slots tested and hits into d_counters[0..2]
cudaError_t launch_bvh_batch_which(int n, const gcore_ray* d_rays, float t_near, float t_far, gcore_HitCpp* d_hits, int* d_found, int which,
                                   unsigned long long* d_counters, cudaStream_t s);
// slot-major accumulator -> row-major outputs; either pointer may be null.
// hdr4[y*W+x] = (r,g,b,spp), ldr[y*W+x] = R | G<<8 | B<<16 (ray_tracer.adb:281-291, 19-57)
cudaError_t launch_resolve(const RenderArgs& a, int spp, float4* d_hdr4, uint32_t* d_ldr, cudaStream_t s);
// gather variant for multi-GPU: `all` holds world slot-major buffers back to back (rank r at r*nslots)
cudaError_t launch_resolve_gathered(const RenderArgs& a, const float4* d_all, int spp, float4* d_hdr4, uint32_t* d_ldr, cudaStream_t s);
int render_kernel_regs(int render_type);
// row-major HDR frame (device) -> this rank's slot-major accumulator (checkpoint / resume)
cudaError_t launch_load_accum(const RenderArgs& a, const float4* d_hdr4, cudaStream_t s);

// ---- wavefront render (render_wave.cuh): persistent grid, one wave of resident blocks ----
struct WaveBuffers {
  float4* colbuf = nullptr;              // [units][nslots] colour of every (pass, pixel) work item
  size_t  colbuf_items = 0;
  float4* stackA = nullptr;              // [max_depth][wave_total_slots] per-bounce (explicit.xyz, weight.x)
  float2* stackB = nullptr;              //                                                (weight.y, weight.z)
  unsigned int* next_item = nullptr;     // global work counter
  float4* tailbuf = nullptr;             // [tail_items][4] radiance of the single-path work items at the end of a launch
  int     tail_items = 0;                // (pass, pixel) items per launch that are handed out path by path
  int     sm_count = 0;
};
int wave_total_slots(int sm_count);      // path slots resident on the device (grid * slots per block)
int wave_max_depth();
// `units` DoPass units over this rank's slots (units * nslots < 2^31): wavefront kernel + fold into a.accum
// count_mesh: statistics level 2 (counters[1] = rays entering the reference's mesh bbox; otherwise = rays that ran the mesh stage)
cudaError_t launch_render_wave(int render_type, int mesh_kind, const RenderArgs& a, const WaveBuffers& wb, int units, int count_mesh, cudaStream_t s);

// ---- lbvh_build.cu ----
#ifndef LBVH_NODE_BYTES
#define LBVH_NODE_BYTES 32                // packed node: see lbvh_traverse.cuh (64: the full-precision experiment layout)
#endif
struct LbvhBuffers {                     // device allocations owned by the builder
  float4* nodes = nullptr; float4* tris = nullptr; float* tri_normals = nullptr; int* tri_mat = nullptr;
  int ntris = 0, nnodes = 0, root = -1;
  size_t bytes = 0;
  float build_ms = 0.0f;
};
// verts: device float3 array (world space), tris: device int3 array, normals9: optional device per-triangle nA,nB,nC,
// tri_mat: optional device per-triangle material id.  On failure `out` is left empty (nothing half-built).
cudaError_t lbvh_build(const float* d_verts, int nverts, const int* d_tris, int ntris, const float* d_normals9, const int* d_tri_mat,
                       LbvhBuffers* out, cudaStream_t s);
void lbvh_free(LbvhBuffers* b);

}  // namespace gcore_dev
