// render_wave.cuh -- the render hot path as a wavefront over persistent path slots (K1+K2+K4+K5+K6a of SURVEY 2b).
//
// Replaces ray_tracer-integrators.adb:25-301 (DoPass + the three PathTrace variants), ray_tracer.adb:61-132 (camera rays,
// Compute_Shadow) and the Ada task pool of ray_tracer.adb:142-194.
//
// Why this shape.  A one-thread-per-pixel megakernel of this path runs with 12.7 of 32 lanes active and spends 47 % of
// its stall samples waiting for instructions (profiles/r01_render.md): lanes of a warp sit on different
// materials and bounce depths, the 8-triangle mesh loop runs for ~2 lanes, and 60 KB of inlined code thrash the 32 KB
// instruction cache.  Here a persistent block owns NS path slots whose state lives in shared memory (SoA) and walks them
// through stages; between stages slots are re-packed into dense queues, so every stage runs with full warps on one small
// piece of code:
//
//   X   extend       all slots with a resolved path ray : mesh candidate merged with spheres, Cornell box, flat light
//                    (scene.adb:56-86); miss / emitter -> the path ends (queue R); surface -> Philox draws, Lights.Sample,
//                    shadow ray (ray_tracer.adb:100-116), bucket by material type
//   Y   shade        surface slots sorted by material : shadow spheres/box + in_shadow rule (ray_tracer.adb:117-128),
//                    EvalBxDF/EvalPDF + MIS (integrators.adb:161-178, 251-287), SampleAndEvalBxDF (:116,183,291), next
//                    ray;  in the same interval the spare threads take queue R: unwind the finished path
//                    (integrators.adb:124,191,299), add its radiance to the pixel-pass colour in the reference's order
//                    (:42-51), publish the colour when the pass is complete, fetch the next work item, camera ray
//                    (ray_tracer.adb:61-97)
//
// and the mesh queries in between, in one of two ways:
//
//   render_wave_kernel  meshes in constant memory (<= 64 triangles, the reference scenes): block-synchronous, two narrow stages
//                       M1 (shadow rays) / M2 (path rays) of (ray, triangle) pairs, one triangle test per thread
//                       (geometry.adb:231-263), four __syncthreads() per bounce, 1024 slots x 1024 threads per SM.
//   render_bvh_kernel   meshes behind the LBVH: traversal is a chain of dependent memory fetches, so it is taken OUT of the
//                       barrier-synchronous stages.  Half of the block's warps are persistent traversal warps: every lane keeps
//                       one ray's walk in registers (short stack in shared memory), pops the next pending ray from a
//                       shared-memory ring the moment its own finishes (ballot over idle lanes + one CAS per warp), and never
//                       meets a barrier; the other half runs X / Y over the slots whose rays are resolved, synchronised by a
//                       named barrier of their own.  Slots with a ray in the tree simply sit the shading rounds out.
//
// A work item is one DoPass of one pixel (4 fixed sub-pixel paths, or 1): items are handed out from a global counter,
// so no slot idles while others finish long paths, and the colour of item (pass k, pixel q) goes to colbuf[k][q].
// fold_kernel then adds the passes into the accumulator in pass order -- the float additions happen in exactly the
// reference's order (colBuff(x,y) := color + colBuff(x,y), integrators.adb:51), so the result is bit-identical to the
// oracle although paths complete out of order.  The recursion's per-bounce terms (explicit light, BSDF weight) are
// kept on a per-slot stack in global memory (L2-resident) and unwound innermost-first like the recursion returns.
#pragma once

#include "art_shade.cuh"

namespace art {

// Block shape of render_wave_kernel: NT threads = NT path slots per block, 1024 / NT blocks per SM.
//   1024 x 1 -- the biggest slot pool wins (measured 256x4 105 ms, 384x3 93, 512x2 85.5, 1024x1 77.4)
#define WAVE_TOTAL_THREADS 1024
#define WAVE_MAX_SLOTS_PER_SM 1536         // what the per-slot bounce stacks in global memory are sized for
#ifndef SPH_UNROLL
#define SPH_UNROLL 1
#endif
constexpr int kSphereUnroll = SPH_UNROLL;     // unroll factor of the sphere loop of IntersectAllSpheres
// path slots of a block of NT threads.  Stage Y handles surface slots + regenerated slots (<= NS) plus at most 31 lanes
// of padding between the two groups; with one block per SM nothing hides a second, nearly empty pass over the block, so
// the 1024-thread shape keeps NS + 31 <= NT.  Smaller blocks overlap each other instead and use every thread.
#ifndef WAVE_RESERVE
#define WAVE_RESERVE 0
#endif
#define WAVE_SLOTS(NT) ((NT) >= 1024 ? (NT) - WAVE_RESERVE : (NT))
#define WAVE_MAX_DEPTH 16

#define WAVE_PAIR_TRIS 8                  // constant-memory meshes up to this size are tested as (ray, triangle) pairs

// render_bvh_kernel: 512 slots, 288 shading + 288 traversal threads per block, two blocks per SM (measured on C5 at 960x540:
// 256+256 23.4 ms, 288+288 22.4, 320+192 27.1, 192+320 24.7, 128+384 28.3, 384+128 38.1; one block of 1024 slots 23.5;
// refill every 2 / 4 / 8 / 16 nodes 24.0 / 23.2 / 23.0 / 23.8)
#ifndef BVH_SLOTS
#define BVH_SLOTS 512
#endif
#ifndef BVH_SHADE_THREADS
#define BVH_SHADE_THREADS 288
#endif
#ifndef BVH_TRAV_THREADS
#define BVH_TRAV_THREADS 288
#endif
#ifndef BVH_BLOCKS_PER_SM
#define BVH_BLOCKS_PER_SM (WAVE_TOTAL_THREADS / BVH_SLOTS)      // 1024 path slots per SM, like render_wave_kernel
#endif
#ifndef BVH_POSTPONE_LEAVES
#define BVH_POSTPONE_LEAVES 1             // traversal warps test parked leaves together (lbvh_traverse.cuh: bvh_walk_step)
#endif
#ifndef BVH_LEAF_LANES
#define BVH_LEAF_LANES 6                  // lanes holding a parked leaf that end a burst of inner-node steps
#endif
#ifndef BVH_REFILL_STEPS
#define BVH_REFILL_STEPS 8                // nodes a traversal lane walks between two looks at the ring of pending rays
#endif

// material bucket of a surface slot = its place in stage Y's thread order (measured on the bench workload: Lambert, Phong, glass,
// mirror 77.8 ms; Phong first 83.4 ms)
#if defined(BUCKET_ORDER_PHONG_FIRST)
#define BUCKET_OF(type) ((type) == MAT_PHONG ? 0 : (type) == MAT_GLASS ? 1 : (type) == MAT_LAMBERT ? 2 : (type) == MAT_MIRROR ? 3 : 4)
#elif defined(BUCKET_ORDER_PHONG_LAST)
#define BUCKET_OF(type) ((type) == MAT_LAMBERT ? 0 : (type) == MAT_GLASS ? 1 : (type) == MAT_MIRROR ? 2 : (type) == MAT_PHONG ? 4 : 3)
#else
#define BUCKET_OF(type) ((type) == MAT_LAMBERT ? 0 : (type) == MAT_PHONG ? 1 : (type) == MAT_GLASS ? 2 : (type) == MAT_MIRROR ? 3 : 4)
#endif
enum { ST_DEPTH_MASK = 31, ST_SUB_SHIFT = 5, ST_SPEC = 1 << 7, ST_HASRAY = 1 << 8, ST_FOLD = 1 << 9, ST_SURF = 1 << 10,
       ST_BUCKET_SHIFT = 11 /* 3 bits: material bucket of a surface slot */ };

// per-slot path state shared by both kernels (SoA, NS slots)
template <int NS>
struct WavePathSmem {
  // path ray
  float ox[NS], oy[NS], oz[NS], dx[NS], dy[NS], dz[NS];
  // surface hit + light sample of the current bounce (X -> Y)
  float ht[NS], nx[NS], ny[NS], nz[NS];
  float sx[NS], sy[NS], sz[NS], ls_pdf[NS], sh_thi[NS];
  float r1[NS], r2[NS];
  float prev_pdf[NS];
  float ix[NS], iy[NS], iz[NS];       // 1 / direction of the slot's pending ray: the path ray (Y -> X), then the shadow ray (X -> Y)
  // colour of the work item so far / radiance at the end of the path
  float cr[NS], cg[NS], cb[NS], lx[NS], ly[NS], lz[NS];
  int   item[NS];                    // unit * nslots + q (the colbuf index); -1: no item; <= -2: a single-path tail item, tailbuf index -2 - item
  unsigned int pixel[NS], sample[NS];   // Philox key word 1 / counter word 0 of the current path
  int   state[NS];                   // depth | sub << 5 | flags
  short hmat[NS];
  short q_r[2][NS];                  // slots whose path ended (or that starve): regenerate queue, double-buffered
  int   cnt_r[2];
  int   pool_next, pool_end, exhausted, retired;
};

// render_wave_kernel: + constant-memory mesh stage buffers and the material sort
template <int NT>
struct WaveSmem : WavePathSmem<NT> {
  float mesh_res[WAVE_PAIR_TRIS * NT];  // mesh stage results per candidate (see wave_mesh_stage)
  short cand[NT];                    // candidate index of the slot's pending ray in the mesh stage, -1: none
  short mrank[NT];                   // position of a surface slot inside its material bucket (counting sort X -> M1 -> Y)
  short q_cs[NT], q_cp[NT], q_s[NT];
  int   cnt_cs, cnt_cp, cnt_m[5];
};

// render_bvh_kernel: + mesh query results, the ring of pending rays and the traversal lanes' short stacks
struct BvhWaveSmem : WavePathSmem<BVH_SLOTS> {
  float mesh_res[4 * BVH_SLOTS];     // result of the slot's last mesh query: t, u, v, bits(triangle id or -1)
  short pend[BVH_SLOTS];             // 0: no ray in the tree; 1 = path ray, 2 = shadow ray queued / being walked (written back to 0 by the traversal lane)
  short tq[BVH_SLOTS];               // ring of slots with a pending ray: producers = shading threads, consumers = traversal warps
  short xq[BVH_SLOTS];               // ring of slots whose path ray is resolved (ready for X): producers = traversal lanes and Y, -1 = not written yet
  short q_sb[5][BVH_SLOTS];          // surface slots whose shadow query is resolved, per material bucket (consumed by Y)
  int   tstack[LBVH_SHORT_STACK * BVH_TRAV_THREADS];   // short stack of traversal lane g: entry k at [k * BVH_TRAV_THREADS + g]
  int   tq_head, tq_tail, tq_pub;    // consumed / reserved by producers / published (entries below tq_pub are written)
  int   xq_tail, xq_hi;              // reserved by producers / snapshot taken for this round's X
  int   cnt_m[5];
  int   done, xwork[2];              // xwork[parity]: did this round's X find a slot to extend?
};

// warp-aggregated append of `slot` to a shared-memory queue (one atomic per warp); returns the position
ART_D int queue_push(bool pred, short* q, int* cnt, int slot) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return -1;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cnt, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  const int pos = base + __popc(m & ((1u << lane) - 1u));
  if (pred) q[pos] = (short)slot;
  return pred ? pos : -1;
}

// the same into a ring of `size` entries with a free-running reservation counter
ART_D void ring_push(bool pred, short* q, int* tail, int size, int slot) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(tail, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) q[(unsigned)(base + __popc(m & ((1u << lane) - 1u))) % (unsigned)size] = (short)slot;
}

// rank `slot` inside the material bucket of its warp peers and file it in that bucket's queue (the counting sort of the surface
// slots, done at the moment a slot is known to be ready for Y).  All 32 lanes call; pred selects.
ART_D void bucket_push(bool pred, int bucket, int slot, int* cnt5, short* q, int stride) {
  const int lane = threadIdx.x & 31;
  const unsigned m = __match_any_sync(0xffffffffu, pred ? bucket : (-1 - lane));
  if (pred) {
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cnt5[bucket], __popc(m));
    base = __shfl_sync(m, base, leader);
    q[bucket * stride + base + __popc(m & ((1u << lane) - 1u))] = (short)slot;
  }
}

struct WaveArgs {
  gcore_dev::RenderArgs r;
  float4* colbuf;                         // [units][nslots]
  float4* stackA;                         // [WAVE_MAX_DEPTH][total_slots]  (E.xyz, W.x)
  float2* stackB;                         // [WAVE_MAX_DEPTH][total_slots]  (W.y, W.z)
  unsigned int* next_item;                // global work counter
  float4* tailbuf;                        // [total_items - tail_first][4] radiance of the single-path items of the launch's tail
  unsigned int total_items;               // units * nslots: the (pass, pixel) work items of this launch
  unsigned int tail_first;                // hand-out positions >= tail_first go out as 4 single sub-pixel paths each (== total_items: none)
  int     units;
  int     count_mesh;                     // statistics level 2: count reference-bbox entries (see mesh_candidate)
};

// does the ray need the mesh stage?  count: also count the rays that enter the reference's mesh bbox (the 564-flop queries
// of SURVEY 8d) -- a statistic that needs the bbox test for every ray; without it the cheap cull goes first and the
// reference's test runs for the few rays that survive it (the decision is the same: both tests must pass)
template <int MESHK>
ART_D bool mesh_candidate(f3 o, f3 inv, bool count, unsigned int& mesh_queries) {
  if (MESHK == MESHK_CONST) { if (g_scn.ntris_bf <= 0) return false; }
  else { if (g_scn.bvh.ntris <= 0) return false; }
  const bool ref_box = (MESHK == MESHK_CONST) && (g_scn.mesh_mode == MESH_REFERENCE_BF);   // IntersectMeshBF bbox reject, geometry.adb:272-274
  float t0, t1;
  if (count) {
    if (ref_box && !intersect_box_inv(o, inv, g_scn.mesh_bmin, g_scn.mesh_bmax, t0, t1)) return false;
    mesh_queries++;
    return cull_box_pass(o, inv, g_scn.mesh_tight_min, g_scn.mesh_tight_max);
  }
  // the reference's bbox also spans the origin (never reset, geometry.adb:593-603): rays that miss the padded tight box
  // of the vertices cannot pass any triangle test
  if (!cull_box_pass(o, inv, g_scn.mesh_tight_min, g_scn.mesh_tight_max)) return false;
  if (ref_box && !intersect_box_inv(o, inv, g_scn.mesh_bmin, g_scn.mesh_bmax, t0, t1)) return false;
  mesh_queries++;
  return true;
}

template <int MESHK>
ART_D void mesh_query(f3 o, f3 d, float& t_o, float& u_o, float& v_o, int& id_o) {
  id_o = -1; t_o = 0.0f; u_o = 0.0f; v_o = 0.0f;
  if (MESHK == MESHK_CONST) {
    const bool ref_mode = (g_scn.mesh_mode == MESH_REFERENCE_BF);
    // every triangle in index order against the window (nearest.tmin, nearest.tmax); in reference mode a hit moves the
    // window to (t, t + 1e-6) (geometry.adb:289-293, 255-256), in nearest mode to (0, t)
    float w_min = 0.0f, w_max = 1000000.0f;
#pragma unroll 1
    for (int i = 0; i < g_scn.ntris_bf; i++) {
      float t, u, v;
      if (intersect_triangle(o, d, ld3(g_scn.triA[i]), ld3(g_scn.triE1[i]), ld3(g_scn.triE2[i]), w_min, w_max, t, u, v)) {
        if (ref_mode) { w_min = t; w_max = t + 1.0e-6f; } else { w_max = t; }
        t_o = t; u_o = u; v_o = v; id_o = i;
      }
    }
  } else {
    const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    bvh_closest(g_scn.bvh, o, d, inv, 0.0f, 1000000.0f, id_o, t_o, u_o, v_o);
  }
}

// Scene.Find_Closest_Hit (scene.adb:56-86) given the mesh candidate: spheres, Cornell box, flat light in the reference's
// order, strict '<'; the record is materialised for the winner only (shadow queries read is_hit / t alone)
template <int MESHK>
ART_D void closest_with_mesh(f3 o, f3 d, f3 inv, int tri_id, float tri_t, float tri_u, float tri_v, bool want_record, DHit& h) {
  float min_t = ART_FLT_LAST;                                                 // IntersectAllSpheres, geometry.adb:48-115
  int   min_i = 0;
#pragma unroll kSphereUnroll
  for (int i = 0; i < g_scn.nspheres; i++) {
    f3 k = mk3(o.x - g_scn.sph[i][0], o.y - g_scn.sph[i][1], o.z - g_scn.sph[i][2]);
    float b = dot(k, d);
    float c = dot(k, k) - g_scn.sph[i][3] * g_scn.sph[i][3];
    float dd = b * b - c;
    if (dd >= 0.0f) {
      float sqrtd = sqrtf(dd);
      float t1 = -b - sqrtd;
      float t2 = -b + sqrtd;
      if (t1 > 0.0f && t1 < min_t) { min_t = t1; min_i = i; }
      else if (t2 > 0.0f && t2 < min_t) { min_t = t2; min_i = i; }
    }
  }
  const bool sph_hit = (min_t > 0.0f && min_t < ART_FLT_LAST);
  int   winner = 0;
  float nearestHitDist = ART_FLT_LAST;
  if (sph_hit && min_t < nearestHitDist) { winner = 0; nearestHitDist = min_t; }

  int box_plane = 0;                                                          // IntersectCornellBox, geometry.adb:193-229
  {
    float tmin, tmax;
    if (intersect_box_inv(o, inv, g_scn.box_min, g_scn.box_max, tmin, tmax)) {
      if (tmax < nearestHitDist) {
        f3 p = mk3(o.x + tmax * d.x, o.y + tmax * d.y, o.z + tmax * d.z);
        const float eps = 1.0e-5f;
        int planeId = 0;
        if (fabsf(p.x - g_scn.box_min[0]) < eps) planeId = 0;
        if (fabsf(p.x - g_scn.box_max[0]) < eps) planeId = 1;
        if (fabsf(p.y - g_scn.box_min[1]) < eps) planeId = 2;
        if (fabsf(p.y - g_scn.box_max[1]) < eps) planeId = 3;
        if (fabsf(p.z - g_scn.box_min[2]) < eps) planeId = 4;
        if (fabsf(p.z - g_scn.box_max[2]) < eps) planeId = 5;
        if (planeId != 5) { winner = 1; nearestHitDist = tmax; box_plane = planeId; }
      }
    }
  }
  if (g_scn.light_shape == LIGHT_RECT) {                                      // IntersectFlatLight, geometry.adb:118-143
    float tmin = (g_scn.fl_max[1] - o.y) * inv.y;
    f3 hp = mk3(o.x + tmin * d.x, o.y + tmin * d.y, o.z + tmin * d.z);
    bool is_hit = (hp.x > g_scn.fl_min[0]) && (hp.x < g_scn.fl_max[0]) && (hp.z > g_scn.fl_min[2]) && (hp.z < g_scn.fl_max[2]) && (tmin >= 0.0f);
    if (is_hit && tmin < nearestHitDist) { winner = 2; nearestHitDist = tmin; }
  }
  if (tri_id >= 0 && tri_t < nearestHitDist) { winner = 3; nearestHitDist = tri_t; }   // IntersectMeshBF result (mesh stages M2 / M1)

  h.is_hit = (winner == 0) ? (sph_hit ? 1 : 0) : 1;
  h.t = (winner == 0) ? (sph_hit ? min_t : 1.0f) : nearestHitDist;
  if (!want_record) return;
  h.tx = 0.0f; h.ty = 0.0f;
  if (winner == 0) {
    h.prim_type = PRIM_SPHERE; h.prim_index = min_i; h.matId = 0;
    h.mat = (g_scn.nspheres > 0) ? g_scn.sph_mat[min_i] : -1;
    if (sph_hit) {
      f3 p = mk3(o.x + min_t * d.x, o.y + min_t * d.y, o.z + min_t * d.z);
      h.normal = normalize(mk3(p.x - g_scn.sph[min_i][0], p.y - g_scn.sph[min_i][1], p.z - g_scn.sph[min_i][2]));
    } else { h.normal = mk3(0.0f, 1.0f, 0.0f); }
  } else if (winner == 1) {
    h.prim_type = PRIM_PLANE; h.prim_index = box_plane;
    h.matId = g_scn.box_mat[box_plane]; h.mat = -1;
    h.normal = ld3(&g_scn.box_nrm[3 * box_plane]);
  } else if (winner == 2) {
    h.prim_type = PRIM_QUAD; h.prim_index = 0; h.matId = 0; h.mat = 4;
    h.normal = mk3(0.0f, -1.0f, 0.0f);
  } else {
    h.prim_type = PRIM_TRIANGLE; h.prim_index = tri_id; h.mat = -1;
    h.matId = (MESHK == MESHK_BVH && g_scn.bvh.tri_mat != nullptr) ? __ldg(g_scn.bvh.tri_mat + tri_id) : 2;   // geometry.adb:312 / gcore_set_mesh_material
    const float* n9 = (MESHK == MESHK_CONST) ? g_scn.triN[tri_id] : (g_scn.bvh.tri_normals + 9 * (size_t)tri_id);
    float w = 1.0f - tri_u - tri_v;                                            // geometry.adb:301
    h.normal = mk3(w * n9[0] + tri_v * n9[3] + tri_u * n9[6],
                   w * n9[1] + tri_v * n9[4] + tri_u * n9[7],
                   w * n9[2] + tri_v * n9[5] + tri_u * n9[8]);
    h.tx = w * 0.0f + tri_v * 0.0f + tri_u * 0.0f;                             // uvs zeroed by the loader, geometry.adb:565-566
    h.ty = h.tx;
  }
  if (h.mat < 0) h.mat = h.matId;                                              // scene.adb:80-82
}

// slot q of this rank -> pixel (tile layout shared with the resolve kernel, see render_kernels.cu)
struct WavePixel { int x, y, active; };
ART_D WavePixel wave_slot_to_pixel(const gcore_dev::RenderArgs& a, int q) {
  int tile_local = q >> 10, r = q & 1023;
  int tile = tile_local * a.world + a.rank;
  int tx, ty; gcore_dev::gc_tile_to_xy(tile, a.tiles_x, tx, ty);
  int w = r >> 5, l = r & 31;
  WavePixel p;
  p.x = tx * GC_TILE + (w & 3) * 8 + (l & 7);
  p.y = ty * GC_TILE + (w >> 2) * 4 + (l >> 3);
  p.active = (tile < a.ntiles) && (p.x < a.width) && (p.y < a.height);
  return p;
}

ART_D f3 wave_eye_ray_dir(const gcore_dev::RenderArgs& a, int x, int y, int sub) {   // ray_tracer.adb:61-97, integrators.adb:46,58
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

// raw Moeller-Trumbore values of geometry.adb:238-248 (no acceptance test): the same operation sequence as
// intersect_triangle, used to re-derive (t, u, v) of the triangle the pair stage selected
ART_D void triangle_tuv(f3 o, f3 d, f3 A, f3 edge1, f3 edge2, float& t, float& u, float& v) {
  f3 pvec = cross(d, edge2);
  f3 tvec = o - A;
  float invDet = 1.0f / max2(dot(edge1, pvec), 1.0e-25f);
  v = dot(tvec, pvec) * invDet;
  f3 qvec = cross(tvec, edge1);
  u = dot(qvec, d) * invDet;
  t = dot(edge2, qvec) * invDet;
}

// ---------------------------------------------------------------------------------------------------------------------
// the three per-slot pieces of the integrators, shared by both kernels (S: any shared-memory struct built on WavePathSmem)
// ---------------------------------------------------------------------------------------------------------------------

// The order in which the (pass, pixel) work items of a launch are handed out.  Position p of the work counter -> (unit, slot q):
// tile-major -- all passes of one 32x32 tile (1024 consecutive slots) before the next tile -- so that the paths in flight on a GPU
// at any time belong to a few tiles: coherent hit types and materials in the warps of X and Y.  (Pass-major order keeps a band of
// the image in flight on one GPU but the WHOLE share of a GPU when 8 of them split the tiles; measured 14.4 vs 15.4 Grays/s per GPU.)
// The colour of an item still goes to colbuf[unit * nslots + q], so fold_kernel's pass order is untouched.
struct WaveItem { int unit, q; };
ART_D WaveItem wave_item_at(unsigned int p, int units, int nslots) {
  WaveItem it;
#ifdef WAVE_ITEM_ORDER_PASS_MAJOR
  it.unit = (int)(p / (unsigned int)nslots); it.q = (int)(p - (unsigned int)it.unit * (unsigned int)nslots);
#elif defined(WAVE_ITEM_ORDER_BLOCK_MAJOR)
  const unsigned int per_blk = (unsigned int)units << 5;                            // all passes of one 8x4 pixel block before the next block
  const unsigned int b = p / per_blk, r = p - b * per_blk;
  it.unit = (int)(r >> 5); it.q = (int)((b << 5) + (r & 31u));
#else
  const unsigned int per_tile = (unsigned int)units << 10;
  const unsigned int tl = p / per_tile, r = p - tl * per_tile;
  it.unit = (int)(r >> 10); it.q = (int)((tl << 10) + (r & 1023u));
#endif
  return it;
}
ART_D unsigned int wave_item_position(int unit, int q, int units, int nslots) {     // the inverse: where (unit, q) sits in the hand-out order
#ifdef WAVE_ITEM_ORDER_PASS_MAJOR
  return (unsigned int)unit * (unsigned int)nslots + (unsigned int)q;
#elif defined(WAVE_ITEM_ORDER_BLOCK_MAJOR)
  return ((unsigned int)(q >> 5) * (unsigned int)units + (unsigned int)unit) * 32u + (unsigned int)(q & 31);
#else
  return ((unsigned int)(q >> 10) * (unsigned int)units + (unsigned int)unit) * 1024u + (unsigned int)(q & 1023);
#endif
}

struct ExtendOut { bool ended, surf, push_cs; int bucket; };

// stage X for slot s whose path ray is resolved: (tri_*) = its mesh candidate.  Scene.Find_Closest_Hit, then path end
// (miss / emitter: integrators.adb:98-108, 151-157, 218-247) or, on a surface, the Philox draws, Lights.Sample and the shadow
// ray of Compute_Shadow (ray_tracer.adb:100-116).  Writes the slot's state word.
template <int RT, int MESHK, class SM>
ART_D ExtendOut wave_extend(SM& S, int s, int st, const gcore_dev::RenderArgs& a, int tri_id, float tri_t, float tri_u, float tri_v,
                            bool cnt, unsigned int& nmesh, unsigned int& nbounce) {
  ExtendOut out; out.ended = false; out.surf = false; out.push_cs = false; out.bucket = -1 - (int)(threadIdx.x & 31);   // distinct per lane: no peers in a match
  const f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
  const f3 inv = mk3(S.ix[s], S.iy[s], S.iz[s]);                        // 1.0/direction (geometry.adb:154-156), computed when the ray was made
  DHit h;
  closest_with_mesh<MESHK>(o, d, inv, tri_id, tri_t, tri_u, tri_v, true, h);
  st &= ~(ST_SURF | (7 << ST_BUCKET_SHIFT));
  f3 Lterm = mk3(0.f, 0.f, 0.f);
  if (!h.is_hit) {                                                      // integrators.adb:98,151,218
    out.ended = true;
  } else if (g_scn.mats[h.mat].type == MAT_LIGHT) {
    out.ended = true;
    if (RT != GCORE_PT_SHADOW && !(dot(neg(d), h.normal) < 0.0f)) {     // :155-157 ; :103, :224
      f3 emit = (g_scn.mats[h.mat].light < 0) ? mk3(0.f, 0.f, 0.f) : ld3(g_scn.l_intensity);   // materials.adb:147-156
      if (RT == GCORE_PT_STUPID) Lterm = emit;                          // :106
      else {                                                            // :230-243
        float lgtPdf = light_eval_pdf(o, d, h.t);
        float bsdfPdf = S.prev_pdf[s];
        float misWeight = (st & ST_SPEC) ? 1.0f : (bsdfPdf * bsdfPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf));
        Lterm = emit * misWeight;
      }
    }
  } else {
    // ---- surface: Philox draws, light sample + shadow ray (Compute_Shadow, ray_tracer.adb:100-116)
    nbounce++;
    out.surf = true;
    st |= ST_SURF;
    const BounceRand br = bounce_rand(a.seed, S.pixel[s], S.sample[s], (uint32_t)(st & ST_DEPTH_MASK));
    S.r1[s] = br.u2; S.r2[s] = br.u3;
    S.ht[s] = h.t; S.nx[s] = h.normal.x; S.ny[s] = h.normal.y; S.nz[s] = h.normal.z; S.hmat[s] = (short)h.mat;
    if (RT != GCORE_PT_STUPID) {                                        // :161-178 / :251-287
      const f3 hpos = mk3(o.x + h.t * d.x, o.y + h.t * d.y, o.z + h.t * d.z);
      LightSample ls = light_sample(br.u0, br.u1, hpos);
      float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
      epsilon = max2(epsilon, 1.0e-30f);
      const f3 sdir = normalize(ls.pos - hpos);
      const f3 so = hpos + sdir * epsilon;
      float maxDist  = length(hpos - ls.pos);
      float epsilon2 = max2(maxDist * 0.000001f, 1.0e-30f);
      S.sx[s] = sdir.x; S.sy[s] = sdir.y; S.sz[s] = sdir.z; S.ls_pdf[s] = ls.pdf; S.sh_thi[s] = maxDist - epsilon2;
      const f3 sinv = mk3(1.0f / sdir.x, 1.0f / sdir.y, 1.0f / sdir.z);
      S.ix[s] = sinv.x; S.iy[s] = sinv.y; S.iz[s] = sinv.z;
      out.push_cs = mesh_candidate<MESHK>(so, sinv, cnt, nmesh);
    }
    const int type = g_scn.mats[h.mat].type;
    out.bucket = BUCKET_OF(type);
    st |= out.bucket << ST_BUCKET_SHIFT;
  }
  if (out.ended) {
    S.lx[s] = Lterm.x; S.ly[s] = Lterm.y; S.lz[s] = Lterm.z;
    st = (st & ~ST_HASRAY) | ST_FOLD;
  }
  S.state[s] = st;
  return out;
}

struct ShadeOut { bool ended, newray, push_cp; };

// stage Y for surface slot s whose shadow query is resolved: (tri_id, tri_t) = the shadow ray's mesh candidate.  The in_shadow
// rule, EvalBxDF / EvalPDF / MIS, SampleAndEvalBxDF, the bounce terms onto the slot's stack, the next path ray.
template <int RT, int MESHK, class SM>
ART_D ShadeOut wave_shade(SM& S, int s, const WaveArgs& wa, size_t stack_index, int total_slots, int tri_id, float tri_t,
                          bool cnt, unsigned int& nmesh, unsigned int& nrays) {
  ShadeOut out; out.ended = false; out.newray = false; out.push_cp = false;
  int st = S.state[s];
  const f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
  const float ht = S.ht[s];
  const f3 hn = mk3(S.nx[s], S.ny[s], S.nz[s]);
  const f3 hpos = mk3(o.x + ht * d.x, o.y + ht * d.y, o.z + ht * d.z);
  const DMaterial& m = g_scn.mats[S.hmat[s]];
  // reflect(ray_dir, normal): Phong eval/sample, mirror, glass (materials.adb:251,313,371,394); Lambert never reads it
  const f3 refl = (m.type == MAT_LAMBERT) ? hn : reflect(d, hn);
  f3 explicitColor = mk3(0.f, 0.f, 0.f);
  if (RT != GCORE_PT_STUPID) {
    const f3 sdir = mk3(S.sx[s], S.sy[s], S.sz[s]);
    float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
    epsilon = max2(epsilon, 1.0e-30f);
    const f3 so = hpos + sdir * epsilon;
    DHit hs;                                                        // closest hit; only is_hit / t are read (ray_tracer.adb:117-122)
    closest_with_mesh<MESHK>(so, sdir, mk3(S.ix[s], S.iy[s], S.iz[s]), tri_id, tri_t, 0.0f, 0.0f, false, hs);
    nrays++;
    const bool shadowed = hs.is_hit && (hs.t < S.sh_thi[s] && hs.t > 10.0f * epsilon);
    if (!shadowed) {
      f3 bxdfVal; float bsdfPdf;
      mat_eval(m, sdir, hn, refl, bxdfVal, bsdfPdf);
      float cosTheta1 = max2(dot(sdir, hn), 0.0f);
      f3 intensity = ld3(g_scn.l_intensity);
      const float ls_pdf = S.ls_pdf[s];
      if (RT == GCORE_PT_SHADOW) {
        explicitColor = intensity * (cosTheta1 * bxdfVal) * (1.0f / max2(ls_pdf, 1.0e-20f));
      } else {
        float lgtPdf = ls_pdf;
        float misWeight = lgtPdf * lgtPdf / (lgtPdf * lgtPdf + bsdfPdf * bsdfPdf);
        explicitColor = intensity * (1.0f / max2(lgtPdf, 1.0e-20f)) * (cosTheta1 * bxdfVal) * misWeight;
      }
    }
  }
  MatSampleD ms = mat_sample(m, S.r1[s], S.r2[s], d, hn, refl);     // pick up next ray, :116,183,291
  f3 bxdfVal = ms.color * (1.0f / max2(ms.pdf, 1.0e-20f));
  float cosTheta = dot(ms.direction, hn);
  const f3 W = fabsf(cosTheta) * bxdfVal;
  int depth = st & ST_DEPTH_MASK;
  wa.stackA[(size_t)depth * total_slots + stack_index] = make_float4(explicitColor.x, explicitColor.y, explicitColor.z, W.x);
  wa.stackB[(size_t)depth * total_slots + stack_index] = make_float2(W.y, W.z);
  depth++;
  const f3 off3 = (sign_ada(cosTheta) * hn) * 1.0e-5f;              // :122,189,297
  const f3 no = hpos + off3, nd = ms.direction;
  S.ox[s] = no.x; S.oy[s] = no.y; S.oz[s] = no.z; S.dx[s] = nd.x; S.dy[s] = nd.y; S.dz[s] = nd.z;
  S.prev_pdf[s] = ms.pdf;
  st = (st & ~(ST_DEPTH_MASK | ST_SPEC | ST_SURF)) | depth | (ms.pureSpecular ? ST_SPEC : 0);
  if (depth == wa.r.max_depth) {                                    // recursion_level = 0 -> (0,0,0)
    S.lx[s] = 0.f; S.ly[s] = 0.f; S.lz[s] = 0.f;
    st = (st & ~ST_HASRAY) | ST_FOLD;
    out.ended = true;
  } else {
    out.newray = true;
    const f3 ninv = mk3(1.0f / nd.x, 1.0f / nd.y, 1.0f / nd.z);
    S.ix[s] = ninv.x; S.iy[s] = ninv.y; S.iz[s] = ninv.z;
    out.push_cp = mesh_candidate<MESHK>(no, ninv, cnt, nmesh);
  }
  S.state[s] = st;
  return out;
}

struct RegenOut { bool starved, retire, newray, push_cp; };

// regenerate for the slot at position r of queue R (on: this lane has one; every lane of the warp calls): unwind the finished
// path innermost-first (integrators.adb:124,191,299), add its radiance to the colour of the pixel pass in the reference's order
// (:42-51), publish a completed pass, fetch the next work item (one atomic per warp), camera ray (ray_tracer.adb:61-97)
template <int RT, int MESHK, class SM>
ART_D RegenOut wave_regen(SM& S, int s, bool on, const WaveArgs& wa, size_t stack_index, int total_slots, bool cnt, unsigned int& nmesh) {
  const gcore_dev::RenderArgs& a = wa.r;
  const int aa = a.anti_aliasing;
  RegenOut out; out.starved = false; out.retire = false; out.newray = false; out.push_cp = false;
  bool need = false, padded = false;
  int st = 0, item = -1;
  if (on) {
    st = S.state[s];
    item = S.item[s];
    if (st & ST_FOLD) {                                               // unwind the recursion: integrators.adb:124,191,299
      f3 L = mk3(S.lx[s], S.ly[s], S.lz[s]);
      const int depth = st & ST_DEPTH_MASK;
#pragma unroll 1
      for (int k0 = depth - 1; k0 >= 0; k0 -= 4) {                    // four levels of loads in flight at a time
        float4 A[4]; float2 B[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int k = (k0 - j >= 0) ? (k0 - j) : 0;
          A[j] = wa.stackA[(size_t)k * total_slots + stack_index];
          B[j] = wa.stackB[(size_t)k * total_slots + stack_index];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (k0 - j >= 0) {
            const f3 W = mk3(A[j].w, B[j].x, B[j].y);
            if (RT == GCORE_PT_STUPID) L = W * L;
            else L = mk3(A[j].x, A[j].y, A[j].z) + W * L;
          }
        }
      }
      const int sub = (st >> ST_SUB_SHIFT) & 3;
      if (item < -1) {
        // an item of the launch's tail: ONE sub-pixel path; its radiance goes out as it is and fold_kernel adds the four in order
        wa.tailbuf[-2 - item] = make_float4(L.x, L.y, L.z, 0.0f);
        item = -1;
        st &= ~ST_FOLD;
      } else {
        f3 color = mk3(S.cr[s], S.cg[s], S.cb[s]);
        bool done;
        if (aa) {
          if (sub == 0) color = mk3(a.bg[0], a.bg[1], a.bg[2]);       // :42
          color = color + L;                                          // :47
          done = (sub == 3);
        } else { color = L; done = true; }                            // :60
        if (done) { wa.colbuf[item] = make_float4(color.x, color.y, color.z, 0.0f); item = -1; }
        else { S.cr[s] = color.x; S.cg[s] = color.y; S.cb[s] = color.z; }
        st = (st & ~(ST_FOLD | (3 << ST_SUB_SHIFT))) | (((sub + 1) & 3) << ST_SUB_SHIFT);
      }
    }
    need = item == -1;
    if (!need) S.sample[s] = S.sample[s] + 1u;                        // next sub-pixel sample of the same pass
  }
  {                                                                   // next work item (pass, pixel): one atomic per warp
    const unsigned nm = __ballot_sync(0xffffffffu, need);
    if (nm) {
      const int lane = threadIdx.x & 31, leader = __ffs(nm) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&S.pool_next, __popc(nm));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) {
        const int got = base + __popc(nm & ((1u << lane) - 1u));
        out.starved = true;
        st &= ~(3 << ST_SUB_SHIFT);
        if (got < S.pool_end) {
          // the counter runs over the (pass, pixel) items first and then, for the last items of the launch, over their four
          // sub-pixel paths one by one: the drain of a launch is then one path long, not four (a slot walks the sub-pixel
          // paths of an item one after the other)
          unsigned int pos = (unsigned int)got; int sub = 0, tail_index = -1;
          if (pos >= wa.tail_first) { const unsigned int t = pos - wa.tail_first; pos = wa.tail_first + (t >> 2); sub = (int)(t & 3u); tail_index = (int)t; }
          const WaveItem it = wave_item_at(pos, wa.units, a.nslots);
          const int q = it.q, unit = it.unit;
          const WavePixel px = wave_slot_to_pixel(a, q);
          if (px.active) {                                            // (an item of a padding pixel is dropped; the slot asks again next round)
            out.starved = false;
            item = (tail_index >= 0) ? (-2 - tail_index) : (unit * a.nslots + q);
            S.pixel[s] = (unsigned int)(px.y * a.width + px.x);
            S.sample[s] = (unsigned int)(a.first_sample + (aa ? unit * 4 + sub : unit));
            st |= sub << ST_SUB_SHIFT;
          } else padded = true;
        }
      }
    }
  }
  if (on) {
    S.item[s] = item;
    if (!out.starved) {
      const int sub = (st >> ST_SUB_SHIFT) & 3;
      const unsigned int pixel = S.pixel[s];
      const int py = (int)(pixel / (unsigned int)a.width), pxx = (int)(pixel - (unsigned int)py * (unsigned int)a.width);
      const f3 o = ld3(g_scn.cam_pos);
      const f3 d = wave_eye_ray_dir(a, pxx, py, aa ? sub : -1);
      S.ox[s] = o.x; S.oy[s] = o.y; S.oz[s] = o.z; S.dx[s] = d.x; S.dy[s] = d.y; S.dz[s] = d.z;
      S.prev_pdf[s] = 1.0f;                                           // StartSample, materials.ads:25
      st = (st & (3 << ST_SUB_SHIFT)) | ST_SPEC | ST_HASRAY;          // depth 0
      out.newray = true;
      const f3 cinv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
      S.ix[s] = cinv.x; S.iy[s] = cinv.y; S.iz[s] = cinv.z;
      out.push_cp = mesh_candidate<MESHK>(o, cinv, cnt, nmesh);
    } else {
      st &= ~(ST_HASRAY | ST_FOLD);
      out.retire = (S.exhausted != 0) && !padded;                     // no work left anywhere: the slot retires
    }
    S.state[s] = st;
  }
  return out;
}

// work items are fetched from the global counter in batches that shrink towards the end of the launch (guided
// self-scheduling): 2 NS while there is plenty, down to 64, so that the blocks run dry together.  Thread 0 only.
template <class SM>
ART_D void wave_refill_pool(SM& S, const WaveArgs& wa, unsigned int ns) {
  const unsigned int total_items = wa.tail_first + 4u * (wa.total_items - wa.tail_first);     // work-counter range (see wave_regen)
  const unsigned int seen = *(volatile unsigned int*)wa.next_item;
  const unsigned int rem = (seen < total_items) ? (total_items - seen) : 0u;
  unsigned int want = rem / (2u * gridDim.x);
  want = want < 64u ? 64u : (want > 2u * ns ? 2u * ns : want);
  const unsigned int base = atomicAdd(wa.next_item, want);
  if (base >= total_items) { S.exhausted = 1; S.pool_next = 0; S.pool_end = 0; }
  else { S.pool_next = (int)base; S.pool_end = (int)min(base + want, total_items); }
}

// statistics: one atomic per warp and counter
ART_D void wave_flush_counters(const gcore_dev::RenderArgs& a, unsigned int nrays, unsigned int nmesh, unsigned int nbounce,
                               unsigned int tnodes, unsigned int ttris, unsigned int ntrav) {
  unsigned int total_rays = __reduce_add_sync(0xffffffffu, nrays);
  unsigned int total_mesh = __reduce_add_sync(0xffffffffu, nmesh);
  unsigned int total_bnc  = __reduce_add_sync(0xffffffffu, nbounce);
  unsigned int tn = __reduce_add_sync(0xffffffffu, tnodes), tt = __reduce_add_sync(0xffffffffu, ttris), tv = __reduce_add_sync(0xffffffffu, ntrav);
  if ((threadIdx.x & 31) == 0) {
    if (total_rays) {
      atomicAdd(&a.counters[0], (unsigned long long)total_rays);
      atomicAdd(&a.counters[1], (unsigned long long)total_mesh);
      atomicAdd(&a.counters[2], (unsigned long long)total_bnc);
    }
    if (tn | tv) {
      atomicAdd(&a.counters[3], (unsigned long long)tn); atomicAdd(&a.counters[4], (unsigned long long)tt); atomicAdd(&a.counters[5], (unsigned long long)tv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// render_wave_kernel: constant-memory meshes, block-synchronous
// ---------------------------------------------------------------------------------------------------------------------

// mesh stage (M1 / M2) for the rays queued in q_c.  SHADOW: the ray is the shadow ray of the slot's hit.  Meshes of up to
// WAVE_PAIR_TRIS triangles are tested as (ray, triangle) pairs, one triangle test per thread, so the stage is as short as one
// triangle test (nothing else runs on the SM meanwhile): mesh_res[c * ntris + i] = t of triangle i if its (u, v) test passes,
// NaN otherwise.  Bigger meshes (<= 64 triangles): one ray per thread, mesh_res[4c..] = t, u, v, id.
template <bool SHADOW, int NT>
ART_D void wave_mesh_stage(WaveSmem<NT>& S, const short* q_c, int n_c, int tid) {
  const int ntris = g_scn.ntris_bf;
  const bool pair_mode = (g_scn.pair_group != 0);
  const int work = pair_mode ? n_c * ntris : n_c;
  for (int j = tid; j < work; j += NT) {
    const int c = pair_mode ? j / ntris : j;
    const int i = pair_mode ? j - c * ntris : 0;
    const int s = q_c[c];
    f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
    if (SHADOW) {                                                             // shadow ray of Compute_Shadow, ray_tracer.adb:109-113
      const float ht = S.ht[s];
      const f3 hpos = mk3(o.x + ht * d.x, o.y + ht * d.y, o.z + ht * d.z);
      d = mk3(S.sx[s], S.sy[s], S.sz[s]);
      float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
      epsilon = max2(epsilon, 1.0e-30f);
      o = hpos + d * epsilon;
    }
    if (pair_mode) {
      float t, u, v;
      triangle_tuv(o, d, ld3(g_scn.triA[i]), ld3(g_scn.triE1[i]), ld3(g_scn.triE2[i]), t, u, v);
      S.mesh_res[j] = (v > 0.0f && u > 0.0f && u + v < 1.0f) ? t : __int_as_float(0x7fc00000);
    } else {
      float t, u, v; int id;
      mesh_query<MESHK_CONST>(o, d, t, u, v, id);
      S.mesh_res[4 * c + 0] = t; S.mesh_res[4 * c + 1] = u; S.mesh_res[4 * c + 2] = v; S.mesh_res[4 * c + 3] = __int_as_float(id);
    }
  }
}

// the mesh candidate of a ray out of mesh_res: the window walk of IntersectMeshBF (geometry.adb:276-296) over the pair
// results in triangle order (u, v of the selected triangle are re-derived when the caller needs them), or the stored
// per-ray result
template <int NT>
ART_D void wave_mesh_result(const WaveSmem<NT>& S, int c, f3 o, f3 d, bool want_uv, int& id, float& t, float& u, float& v) {
  id = -1; t = 0.0f; u = 0.0f; v = 0.0f;
  if (c < 0) return;
  if (g_scn.pair_group != 0) {
    const int ntris = g_scn.ntris_bf;
    const bool ref_mode = (g_scn.mesh_mode == MESH_REFERENCE_BF);
    float w_min = 0.0f, w_max = 1000000.0f;
#pragma unroll 1
    for (int i = 0; i < ntris; i++) {
      const float ti = S.mesh_res[c * ntris + i];
      if (ti > w_min && ti < w_max) {                                         // NaN (failed u/v test) never passes
        if (ref_mode) { w_min = ti; w_max = ti + 1.0e-6f; } else { w_max = ti; }
        t = ti; id = i;
      }
    }
    if (id >= 0 && want_uv) {
      float t2;
      triangle_tuv(o, d, ld3(g_scn.triA[id]), ld3(g_scn.triE1[id]), ld3(g_scn.triE2[id]), t2, u, v);
    }
  } else {
    t = S.mesh_res[4 * c + 0]; u = S.mesh_res[4 * c + 1]; v = S.mesh_res[4 * c + 2]; id = __float_as_int(S.mesh_res[4 * c + 3]);
  }
}

#ifdef WAVE_BLOCKS_CONST
#define WAVE_BLOCKS_PER_SM(NT) WAVE_BLOCKS_CONST
#else
#define WAVE_BLOCKS_PER_SM(NT) (WAVE_TOTAL_THREADS / (NT))
#endif
template <int RT, int NT>
__global__ void __launch_bounds__(NT, WAVE_BLOCKS_PER_SM(NT)) render_wave_kernel(const __grid_constant__ WaveArgs wa) {
  constexpr int NS = WAVE_SLOTS(NT);        // path slots of this block
  constexpr int MESHK = MESHK_CONST;
  extern __shared__ __align__(16) unsigned char wave_smem_raw[];
  WaveSmem<NT>& S = *reinterpret_cast<WaveSmem<NT>*>(wave_smem_raw);
  const gcore_dev::RenderArgs& a = wa.r;
  const int tid = threadIdx.x;
  const int gslot0 = blockIdx.x * NT;
  const int total_slots = gridDim.x * NT;
  unsigned int nrays = 0, nmesh = 0, nbounce = 0;
  const bool cnt = wa.count_mesh != 0;

  for (int s = tid; s < NS; s += NT) { S.item[s] = -1; S.state[s] = 0; S.q_r[0][s] = (short)s; S.cand[s] = -1; }
  if (tid == 0) {
    S.cnt_r[0] = NS; S.cnt_r[1] = 0; S.cnt_cs = 0; S.cnt_cp = 0; S.retired = 0; S.exhausted = 0;
    for (int b = 0; b < 5; b++) S.cnt_m[b] = 0;
    wave_refill_pool(S, wa, NS);
  }
  __syncthreads();

  int par = 0;
  for (;;) {
    // ================================================================= X: closest hit, path end or light sample
    if (tid == 0) {
      S.cnt_cp = 0;                                                           // consumed by M2 of the previous round
      if (!S.exhausted && S.pool_next >= S.pool_end) wave_refill_pool(S, wa, NS);   // (fetched from in Y only)
    }
    for (int s = tid; s < NS; s += NT) {
      const int st = S.state[s];
      const bool on = (st & ST_HASRAY) != 0;
      ExtendOut x; x.ended = false; x.surf = false; x.push_cs = false; x.bucket = -1 - (tid & 31);
      if (on) {
        const f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
        int tri_id; float tri_t, tri_u, tri_v;
        wave_mesh_result<NT>(S, (int)S.cand[s], o, d, true, tri_id, tri_t, tri_u, tri_v);
        x = wave_extend<RT, MESHK>(S, s, st, a, tri_id, tri_t, tri_u, tri_v, cnt, nmesh, nbounce);
        nrays++;
      }
      queue_push(x.ended, S.q_r[par], &S.cnt_r[par], s);
      if (RT != GCORE_PT_STUPID) {
        const int c = queue_push(x.push_cs, S.q_cs, &S.cnt_cs, s);
        if (on) S.cand[s] = (short)c;                                         // -1 when the shadow ray skips the mesh
      }
      // counting sort by material type: rank inside the bucket
      const unsigned m = __match_any_sync(0xffffffffu, x.bucket);
      if (x.bucket >= 0) {
        const int lane = tid & 31, leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&S.cnt_m[x.bucket], __popc(m));
        base = __shfl_sync(m, base, leader);
        S.mrank[s] = (short)(base + __popc(m & ((1u << lane) - 1u)));
      }
    }
    __syncthreads();

    // ================================================================= M1: scatter of the material sort + mesh stage of the shadow rays
    {
      const int c0 = S.cnt_m[0], c1 = S.cnt_m[1], c2 = S.cnt_m[2], c3 = S.cnt_m[3];
      for (int s = tid; s < NS; s += NT) {
        const int st = S.state[s];
        if ((st & (ST_SURF | ST_HASRAY)) == (ST_SURF | ST_HASRAY)) {
          const int b = (st >> ST_BUCKET_SHIFT) & 7;
          const int off_b = (b == 0) ? 0 : (b == 1) ? c0 : (b == 2) ? c0 + c1 : (b == 3) ? c0 + c1 + c2 : c0 + c1 + c2 + c3;
          S.q_s[off_b + S.mrank[s]] = (short)s;
        }
      }
      if (RT != GCORE_PT_STUPID) wave_mesh_stage<true, NT>(S, S.q_cs, S.cnt_cs, tid);
    }
    __syncthreads();

    // ================================================================= Y: shade (slots sorted by material) + regenerate (queue R)
    {
      if (tid == 0) S.cnt_cs = 0;
      // surface slots sorted by material type (buckets back to back: only the warps on a bucket boundary see two types),
      // one padding gap, then queue R
      const int n_surf = S.cnt_m[0] + S.cnt_m[1] + S.cnt_m[2] + S.cnt_m[3] + S.cnt_m[4];
      const int n_m = (n_surf + 31) & ~31;
      const int n_r = S.cnt_r[par];
      const int n_all = n_m + ((n_r + 31) & ~31);
      for (int i0 = tid; i0 < n_all; i0 += NT) {
        bool ended = false, push_cp = false, starved = false, retire = false, newray = false;
        int s = 0;
#ifdef WAVE_REGEN_FIRST
        const int n_rp = n_all - n_m;
        const int i = (i0 < n_rp) ? (n_m + i0) : (i0 - n_rp);                // queue R on the first threads, the shaded slots behind it
#else
        const int i = i0;
#endif
        if (i < n_m) {
          if (i < n_surf) {
            s = S.q_s[i];
            int tri_id = -1; float tri_t = 0.f, tri_u, tri_v;
            if (RT != GCORE_PT_STUPID) {
              // the shadow ray again (ray_tracer.adb:109-113): the pair results hold t per triangle, the window walk picks the candidate
              const f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
              const float ht = S.ht[s];
              const f3 hpos = mk3(o.x + ht * d.x, o.y + ht * d.y, o.z + ht * d.z);
              const f3 sdir = mk3(S.sx[s], S.sy[s], S.sz[s]);
              float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
              epsilon = max2(epsilon, 1.0e-30f);
              wave_mesh_result<NT>(S, (int)S.cand[s], hpos + sdir * epsilon, sdir, false, tri_id, tri_t, tri_u, tri_v);
            }
            const ShadeOut y = wave_shade<RT, MESHK>(S, s, wa, (size_t)(gslot0 + s), total_slots, tri_id, tri_t, cnt, nmesh, nrays);
            ended = y.ended; newray = y.newray; push_cp = y.push_cp;
          }
        } else {
          const int r = i - n_m;
          const bool on = r < n_r;
          if (on) s = S.q_r[par][r];
          const RegenOut g = wave_regen<RT, MESHK>(S, s, on, wa, (size_t)(gslot0 + s), total_slots, cnt, nmesh);
          starved = g.starved; retire = g.retire; newray = g.newray; push_cp = g.push_cp;
        }
        const int c = queue_push(push_cp, S.q_cp, &S.cnt_cp, s);
        if (newray) S.cand[s] = (short)c;
        // paths that reached Max_Trace_Depth and starved slots wait for the next round's regenerate
        queue_push(ended || (starved && !retire), S.q_r[par ^ 1], &S.cnt_r[par ^ 1], s);
        const unsigned rm = __ballot_sync(0xffffffffu, retire);
        if (rm && (tid & 31) == 0) atomicAdd(&S.retired, __popc(rm));
      }
    }
    __syncthreads();
    if (S.retired >= NS) break;          // uniform: `retired` only changes in Y, which the barrier above closes

    // ================================================================= M2: mesh stage of the new path rays
    if (tid == 0) S.cnt_r[par] = 0;
    if (tid < 5) S.cnt_m[tid] = 0;
    wave_mesh_stage<false, NT>(S, S.q_cp, S.cnt_cp, tid);
    par ^= 1;
    __syncthreads();
  }
  wave_flush_counters(a, nrays, nmesh, nbounce, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------------------------------------
// render_bvh_kernel: meshes behind the LBVH -- shading warps and persistent traversal warps in one block
// ---------------------------------------------------------------------------------------------------------------------

ART_D void shade_group_barrier() { asm volatile("bar.sync 1, %0;" :: "n"(BVH_SHADE_THREADS) : "memory"); }
ART_D int  vload(const int* p)   { return *(const volatile int*)p; }
ART_D int  vload(const short* p) { return (int)*(const volatile short*)p; }

template <int RT>
__global__ void __launch_bounds__(BVH_SHADE_THREADS + BVH_TRAV_THREADS, BVH_BLOCKS_PER_SM)
render_bvh_kernel(const __grid_constant__ WaveArgs wa) {
  constexpr int NS = BVH_SLOTS, NSH = BVH_SHADE_THREADS, NTR = BVH_TRAV_THREADS;
  constexpr int MESHK = MESHK_BVH;
  static_assert(NS * BVH_BLOCKS_PER_SM <= WAVE_MAX_SLOTS_PER_SM, "the per-slot stacks are sized for WAVE_MAX_SLOTS_PER_SM slots");
  extern __shared__ __align__(16) unsigned char wave_smem_raw[];
  BvhWaveSmem& S = *reinterpret_cast<BvhWaveSmem*>(wave_smem_raw);
  const gcore_dev::RenderArgs& a = wa.r;
  const int tid = threadIdx.x;
  const int gslot0 = blockIdx.x * NS;
  const int total_slots = gridDim.x * NS;
  unsigned int nrays = 0, nmesh = 0, nbounce = 0, ntrav = 0;
  BvhStats tstat; tstat.nodes = 0; tstat.tris = 0;
  const bool cnt = wa.count_mesh != 0;

  for (int s = tid; s < NS; s += NSH + NTR) { S.item[s] = -1; S.state[s] = 0; S.q_r[0][s] = (short)s; S.pend[s] = 0; S.xq[s] = -1; }
  if (tid == 0) {
    S.cnt_r[0] = NS; S.cnt_r[1] = 0; S.retired = 0; S.exhausted = 0;
    S.tq_head = 0; S.tq_tail = 0; S.tq_pub = 0; S.xq_tail = 0; S.xq_hi = 0; S.done = 0; S.xwork[0] = 0; S.xwork[1] = 0;
    for (int b = 0; b < 5; b++) S.cnt_m[b] = 0;
    wave_refill_pool(S, wa, NS);
  }
  __syncthreads();

  if (tid >= NSH) {
    // ================================================================= traversal warps: one ray per lane, refilled from the ring
    const int g = tid - NSH, lane = tid & 31;
    int* my_stack = &S.tstack[g];
    int ovf[LBVH_OVERFLOW];                                                   // stack entries beyond the shared short stack (rarely touched)
    bool active = false;
    int s = 0, p = 0;
    f3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f), inv = mk3(0.f, 0.f, 0.f);
    BvhRayState c; c.node = -1; c.sp = 0; c.best_t = 0.f; c.best_prim = -1; c.best_u = 0.f; c.best_v = 0.f;
    int pend0 = 0, pend1 = 0;                                                 // leaves this lane has reached and not tested yet
    for (;;) {
      const unsigned idle_m = __ballot_sync(0xffffffffu, !active);
      if (idle_m) {
        // compaction of the traversal lanes: the idle lanes of the warp take the next pending rays (one CAS per warp)
        int take = 0, h = 0;
        if (lane == 0) {
          for (;;) {
            h = vload(&S.tq_head);
            const int avail = vload(&S.tq_pub) - h;
            take = min(avail, __popc(idle_m));
            if (take <= 0) { take = 0; break; }
            if (atomicCAS(&S.tq_head, h, h + take) == h) break;
          }
        }
        take = __shfl_sync(0xffffffffu, take, 0); h = __shfl_sync(0xffffffffu, h, 0);
        if (take > 0) {
          __threadfence_block();                                              // entries and ray data below tq_pub were written before it was published
          const int r = __popc(idle_m & ((1u << lane) - 1u));
          if (!active && r < take) {
            s = vload(&S.tq[(unsigned)(h + r) % (unsigned)NS]);
            p = vload(&S.pend[s]);
            o = mk3(S.ox[s], S.oy[s], S.oz[s]); d = mk3(S.dx[s], S.dy[s], S.dz[s]);
            float t_clip = 1000000.0f;
            if (p == 2) {                                                     // shadow ray of Compute_Shadow, ray_tracer.adb:109-113
              const float ht = S.ht[s];
              const f3 hpos = mk3(o.x + ht * d.x, o.y + ht * d.y, o.z + ht * d.z);
              d = mk3(S.sx[s], S.sy[s], S.sz[s]);
              float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
              epsilon = max2(epsilon, 1.0e-30f);
              o = hpos + d * epsilon;
              // only a hit with t < maxDist - epsilon2 can change in_shadow (ray_tracer.adb:119-128): a mesh hit beyond it either is
              // not the closest hit or fails that test itself, so the walk is clipped there (the nearest hit below it stays exact)
              t_clip = S.sh_thi[s];
            }
            inv = bvh_slab_inv_of(mk3(S.ix[s], S.iy[s], S.iz[s]));
            bvh_state_begin(g_scn.bvh, t_clip, c);
#if BVH_POSTPONE_LEAVES
            if (c.node == -2) { pend0 = ~0; c.node = -1; }                    // a one-triangle mesh has no inner node: its only leaf is slot 0
#endif
            ntrav++;
            active = true;
          }
        }
      }
      if (!__any_sync(0xffffffffu, active)) {
        if (vload(&S.done)) break;
        __nanosleep(100);
        continue;
      }
#if BVH_POSTPONE_LEAVES
      // ---- inner nodes: lanes holding a leaf wait; the burst ends when enough of them do (or nobody is left walking)
#pragma unroll 1
      for (int k = 0; k < BVH_REFILL_STEPS; k++) {
        if (active && pend0 == 0 && c.node >= 0) {
          tstat.nodes++;
          bvh_walk_step(g_scn.bvh, o, inv, 0.0f, c.best_t, c.node, c.sp, pend0, pend1, my_stack, NTR, ovf);
        }
        const unsigned walk_m = __ballot_sync(0xffffffffu, active && pend0 == 0 && c.node >= 0);
        const unsigned leaf_m = __ballot_sync(0xffffffffu, pend0 != 0);
        if (walk_m == 0u || __popc(leaf_m) >= BVH_LEAF_LANES) break;
      }
      // ---- parked leaves, together
      while (pend0 != 0) {
        int dummy_slot;
        tstat.tris++;
        bvh_leaf_test<false>(g_scn.bvh, ~pend0, 1, o, d, 0.0f, c.best_t, c.best_prim, c.best_u, c.best_v, dummy_slot);
        pend0 = pend1; pend1 = 0;
      }
      if (active && c.node < 0) {
        {
#else
      if (active) {
        const bool fin = bvh_state_advance(g_scn.bvh, o, d, inv, 0.0f, BVH_REFILL_STEPS, c, my_stack, NTR, ovf, tstat);
        if (fin) {
#endif
          S.mesh_res[4 * s + 0] = c.best_t; S.mesh_res[4 * s + 1] = c.best_u; S.mesh_res[4 * s + 2] = c.best_v;
          S.mesh_res[4 * s + 3] = __int_as_float(c.best_prim);
          __threadfence_block();
          *(volatile short*)&S.pend[s] = 0;                                   // the slot's owner sees the ray resolved
          if (p == 1) {                                                       // a path ray: the slot is ready for X
            const int at = atomicAdd(&S.xq_tail, 1);
            *(volatile short*)&S.xq[(unsigned)at % (unsigned)NS] = (short)s;
          }
          active = false;
        }
      }
    }
    wave_flush_counters(a, 0u, 0u, 0u, tstat.nodes, tstat.tris, ntrav);
    return;
  }

  // =================================================================== shading warps: X / rank / Y over the slots whose rays are resolved
  int par = 0, x_lo = 0;
  for (;;) {
    // ---- X: closest hit, path end or light sample + shadow ray
    if (tid == 0 && !S.exhausted && S.pool_next >= S.pool_end) wave_refill_pool(S, wa, NS);
    // the slots whose path ray is resolved come in a ring (filed by the traversal lanes as their walks end, and by Y for rays
    // that need no walk): X runs over a dense list, full warps
    const int x_hi = S.xq_hi;
    for (int i0 = x_lo; i0 < x_hi; i0 += NSH) {
      const int i = i0 + tid;
      const bool on = i < x_hi;
      int s = 0;
      ExtendOut x; x.ended = false; x.surf = false; x.push_cs = false; x.bucket = -1;
      if (on) {
        short* e = &S.xq[(unsigned)i % (unsigned)NS];
        while ((s = vload(e)) < 0) { }                                        // reserved a moment ago, written in a moment
        *(volatile short*)e = -1;
        __threadfence_block();                                                // the traversal lane wrote mesh_res before it filed the slot
        const int st = S.state[s];
        const float tri_t = S.mesh_res[4 * s + 0], tri_u = S.mesh_res[4 * s + 1], tri_v = S.mesh_res[4 * s + 2];
        const int tri_id = __float_as_int(S.mesh_res[4 * s + 3]);
        x = wave_extend<RT, MESHK>(S, s, st, a, tri_id, tri_t, tri_u, tri_v, cnt, nmesh, nbounce);
        nrays++;
        S.xwork[par] = 1;
        if (x.surf) { if (x.push_cs) S.pend[s] = 2; else S.mesh_res[4 * s + 3] = __int_as_float(-1); }
      }
      queue_push(x.ended, S.q_r[par], &S.cnt_r[par], s);
      ring_push(x.surf && x.push_cs, S.tq, &S.tq_tail, NS, s);                // the shadow ray goes to the traversal warps
    }
    x_lo = x_hi;
    shade_group_barrier();
    if (tid == 0) { __threadfence_block(); *(volatile int*)&S.tq_pub = S.tq_tail; }
    // ---- rank: surface slots whose shadow query is resolved, bucketed by material
    for (int s = tid; s < NS; s += NSH) {
      const int st = S.state[s];
      const bool ready = ((st & (ST_SURF | ST_HASRAY)) == (ST_SURF | ST_HASRAY)) && vload(&S.pend[s]) == 0;
      bucket_push(ready, (st >> ST_BUCKET_SHIFT) & 7, s, S.cnt_m, &S.q_sb[0][0], NS);
    }
    shade_group_barrier();
    // ---- Y: shade (sorted by material) + regenerate (queue R)
    {
      const int c0 = S.cnt_m[0], c1 = S.cnt_m[1], c2 = S.cnt_m[2], c3 = S.cnt_m[3], c4 = S.cnt_m[4];
      const int n_surf = c0 + c1 + c2 + c3 + c4;
      const int n_m = (n_surf + 31) & ~31;
      const int n_r = S.cnt_r[par];
      const int n_all = n_m + ((n_r + 31) & ~31);
      if (n_surf == 0 && n_r == 0 && S.xwork[par] == 0) __nanosleep(400);          // every slot is in the tree: leave the issue slots to the traversal warps
      for (int i = tid; i < n_all; i += NSH) {
        bool ended = false, push_cp = false, starved = false, retire = false, newray = false;
        int s = 0;
        if (i < n_m) {
          if (i < n_surf) {
            int b = 0, off = 0;
            if (i >= c0) { b = 1; off = c0; if (i >= c0 + c1) { b = 2; off = c0 + c1; if (i >= c0 + c1 + c2) { b = 3; off = c0 + c1 + c2; if (i >= c0 + c1 + c2 + c3) { b = 4; off = c0 + c1 + c2 + c3; } } } }
            s = S.q_sb[b][i - off];
            __threadfence_block();
            const float tri_t = S.mesh_res[4 * s + 0];
            const int tri_id = __float_as_int(S.mesh_res[4 * s + 3]);
            const ShadeOut y = wave_shade<RT, MESHK>(S, s, wa, (size_t)(gslot0 + s), total_slots, tri_id, tri_t, cnt, nmesh, nrays);
            ended = y.ended; newray = y.newray; push_cp = y.push_cp;
          }
        } else {
          const int r = i - n_m;
          const bool on = r < n_r;
          if (on) s = S.q_r[par][r];
          const RegenOut g = wave_regen<RT, MESHK>(S, s, on, wa, (size_t)(gslot0 + s), total_slots, cnt, nmesh);
          starved = g.starved; retire = g.retire; newray = g.newray; push_cp = g.push_cp;
        }
        if (newray) { if (push_cp) S.pend[s] = 1; else S.mesh_res[4 * s + 3] = __int_as_float(-1); }
        ring_push(newray && push_cp, S.tq, &S.tq_tail, NS, s);                // the new path ray goes to the traversal warps ...
        ring_push(newray && !push_cp, S.xq, &S.xq_tail, NS, s);               // ... or, with no mesh candidate, straight to the next X
        // paths that reached Max_Trace_Depth and starved slots wait for the next round's regenerate
        queue_push(ended || (starved && !retire), S.q_r[par ^ 1], &S.cnt_r[par ^ 1], s);
        const unsigned rm = __ballot_sync(0xffffffffu, retire);
        if (rm && (tid & 31) == 0) atomicAdd(&S.retired, __popc(rm));
      }
    }
    shade_group_barrier();
    if (S.retired >= NS) { if (tid == 0) *(volatile int*)&S.done = 1; break; }   // uniform: `retired` only changes in Y
    if (tid == 0) { __threadfence_block(); *(volatile int*)&S.tq_pub = S.tq_tail; S.cnt_r[par] = 0; S.xwork[par] = 0; S.xq_hi = vload(&S.xq_tail); }
    if (tid < 5) S.cnt_m[tid] = 0;
    par ^= 1;
    shade_group_barrier();                                                    // X reads the snapshot of its ring
  }
  wave_flush_counters(a, nrays, nmesh, nbounce, 0u, 0u, 0u);
}

// K6a: acc[q] := colour(pass k, q) + acc[q] for k = 0 .. units-1 -- the additions of integrators.adb:51/:63 in pass order
__global__ void __launch_bounds__(256) fold_kernel(const __grid_constant__ gcore_dev::RenderArgs a, const float4* __restrict__ colbuf,
                                                   const float4* __restrict__ tailbuf, unsigned int tail_first, int units) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= a.nslots) return;
  const WavePixel px = wave_slot_to_pixel(a, q);
  if (!px.active) return;
  float4 acc4 = a.accum[q];
  f3 acc = mk3(acc4.x, acc4.y, acc4.z);
#pragma unroll 4
  for (int k = 0; k < units; k++) {
    const unsigned int pos = wave_item_position(k, q, units, a.nslots);
    f3 color;
    if (pos < tail_first) {
      const float4 c = __ldcs(&colbuf[(size_t)k * a.nslots + q]);
      color = mk3(c.x, c.y, c.z);
    } else {
      // a tail item: its four sub-pixel paths were traced as separate work items; color := Background_Color, then the four
      // additions of integrators.adb:42-48 in sub-pixel order
      const float4* L = tailbuf + 4 * (size_t)(pos - tail_first);
      color = mk3(a.bg[0], a.bg[1], a.bg[2]);
#pragma unroll
      for (int j = 0; j < 4; j++) { const float4 c = __ldcs(&L[j]); color = color + mk3(c.x, c.y, c.z); }
    }
    acc = color + acc;
  }
  a.accum[q] = make_float4(acc.x, acc.y, acc.z, acc4.w);
}

}  // namespace art
