// render_wave.cuh -- the render hot path as a block-synchronous wavefront (K1+K2+K4+K5+K6a of SURVEY 2b).
//
// Replaces ray_tracer-integrators.adb:25-301 (DoPass + the three PathTrace variants), ray_tracer.adb:61-132 (camera rays,
// Compute_Shadow) and the Ada task pool of ray_tracer.adb:142-194.
//
// Why this shape.  A one-thread-per-pixel megakernel of this path runs with 12.7 of 32 lanes active and spends 47 % of
// its stall samples waiting for instructions (profiles/r01_render.md): lanes of a warp sit on different
// materials and bounce depths, the 8-triangle mesh loop runs for ~2 lanes, and 60 KB of inlined code thrash the 32 KB
// instruction cache.  Here a persistent block owns NS path slots whose state lives in shared memory (SoA), and walks
// all of them through two wide and two narrow stages per bounce, separated by __syncthreads().  Between stages slots
// are re-packed into dense queues, so every stage runs with full warps on one small piece of code that all warps of the
// block execute at the same time:
//
//   X   extend       all slots with a ray : mesh candidate of M2 merged with spheres, Cornell box, flat light
//                    (scene.adb:56-86); miss / emitter -> the path ends (queue R); surface -> Philox draws, Lights.Sample,
//                    shadow ray (ray_tracer.adb:100-116), bucket by material type
//   M1  shadow mesh  (shadow ray, triangle) pairs : one triangle test per thread (geometry.adb:231-263) / LBVH traversal
//   Y   shade        surface slots sorted by material : shadow spheres/box + in_shadow rule (ray_tracer.adb:117-128),
//                    EvalBxDF/EvalPDF + MIS (integrators.adb:161-178, 251-287), SampleAndEvalBxDF (:116,183,291), next
//                    ray;  in the same interval the spare threads take queue R: unwind the finished path
//                    (integrators.adb:124,191,299), add its radiance to the pixel-pass colour in the reference's order
//                    (:42-51), publish the colour when the pass is complete, fetch the next work item, camera ray
//                    (ray_tracer.adb:61-97)
//   M2  path mesh    (path ray, triangle) pairs for the rays made in Y
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

// Block shape: NT threads = NT path slots per block, 1024 / NT blocks per SM (template parameter of the kernel).
//   constant-memory meshes: 1024 x 1 -- the biggest slot pool wins (measured 256x4 105 ms, 384x3 93, 512x2 85.5, 1024x1 77.4)
//   LBVH meshes:            256 x 4 -- the traversal slices are latency-bound, several blocks per SM overlap them with the
//                           other blocks' stages (C5 with 16-node slices: 128x8 1738 ms, 256x4 1535, 512x2 1582)
#define WAVE_TOTAL_THREADS 1024
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

enum { ST_DEPTH_MASK = 31, ST_SUB_SHIFT = 5, ST_SPEC = 1 << 7, ST_HASRAY = 1 << 8, ST_FOLD = 1 << 9, ST_SURF = 1 << 10,
       ST_BUCKET_SHIFT = 11 /* 3 bits: material bucket of a surface slot */, ST_READY = 1 << 14 /* LBVH scenes: ranked for this round's Y */ };

#ifndef WAVE_BVH_SLICE
#define WAVE_BVH_SLICE 16                 // LBVH scenes: nodes a pending ray advances per mesh stage
#endif

// mesh-kind specific part of a block's shared memory
template <int NT, bool BVH> struct WaveMeshSmem;
template <int NT> struct WaveMeshSmem<NT, false> {      // constant-memory meshes
  float mesh_res[WAVE_PAIR_TRIS * NT];  // mesh stage results per candidate (see wave_mesh_stage)
  short cand[NT];                    // candidate index of the slot's pending ray in the mesh stage, -1: none
  short mrank[NT];                   // position of a surface slot inside its material bucket (counting sort X -> M1 -> Y)
  short q_cs[NT], q_cp[NT], q_s[NT];
  int   cnt_cs, cnt_cp, cnt_m[5];
};
template <int NT> struct WaveMeshSmem<NT, true> {       // LBVH meshes: the traversal state of the slot's pending ray lives here, not in a thread
  float mesh_res[4 * NT];            // running best, then the result: t, u, v, bits(triangle id or -1)
  int   tnode[NT];                   // cursor: next node to visit
  short tsp[NT];                     //         stack depth
  short pend[NT];                    // traversal in flight for the slot's 1 = path ray, 2 = shadow ray (+4: not started yet)
  int   tstack[LBVH_SHORT_STACK * NT];   // short stack, entry k of slot s at [k * NT + s]; deeper entries: WaveArgs::tstack_ovf
  short q_t[2][NT];                  // pending traversals: [0] advanced by M1, [1] by M2 (each stage hands its unfinished rays to the other queue)
  short q_sb[5][NT];                 // surface slots whose shadow query is resolved, per material bucket (consumed by Y)
  int   cnt_t[2], cnt_m[2][5];       // cnt_m[parity]: bucket sizes for this round's / the next round's Y
};

template <int NT, bool BVH>
struct WaveSmem : WaveMeshSmem<NT, BVH> {
  // path ray
  float ox[NT], oy[NT], oz[NT], dx[NT], dy[NT], dz[NT];
  // surface hit + light sample of the current bounce (X -> Y)
  float ht[NT], nx[NT], ny[NT], nz[NT];
  float sx[NT], sy[NT], sz[NT], ls_pdf[NT], sh_thi[NT];
  float r1[NT], r2[NT];
  float prev_pdf[NT];
  float ix[NT], iy[NT], iz[NT];       // 1 / direction of the slot's pending ray: the path ray (Y -> X), then the shadow ray (X -> Y)
  // colour of the work item so far / radiance at the end of the path
  float cr[NT], cg[NT], cb[NT], lx[NT], ly[NT], lz[NT];
  int   item[NT];                    // unit * nslots + q, or -1: no item
  unsigned int pixel[NT], sample[NT];   // Philox key word 1 / counter word 0 of the current path
  int   state[NT];                   // depth | sub << 5 | flags
  short hmat[NT];
  short q_r[2][NT];
  int   cnt_r[2];
  int   pool_next, pool_end, exhausted, retired;
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

struct WaveArgs {
  gcore_dev::RenderArgs r;
  float4* colbuf;                         // [units][nslots]
  float4* stackA;                         // [WAVE_MAX_DEPTH][total_slots]  (E.xyz, W.x)
  float2* stackB;                         // [WAVE_MAX_DEPTH][total_slots]  (W.y, W.z)
  unsigned int* next_item;                // global work counter
  int*    tstack_ovf;                     // LBVH scenes: [total_slots][LBVH_OVERFLOW] traversal stack entries beyond the shared short stack
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

// mesh stage (M1 / M2) of constant-memory meshes for the rays queued in q_c.  SHADOW: the ray is the shadow ray of the slot's
// hit.  Meshes of up to WAVE_PAIR_TRIS triangles are tested as (ray, triangle) pairs, one triangle test per thread, so the
// stage is as short as one triangle test (nothing else runs on the SM meanwhile): mesh_res[c * ntris + i] = t of
// triangle i if its (u, v) test passes, NaN otherwise.  Bigger meshes: one ray per thread, mesh_res[4c..] = t, u, v, id.
template <int MESHK, bool SHADOW, int NT>
ART_D void wave_mesh_stage(WaveSmem<NT, false>& S, const short* q_c, int n_c, int tid) {
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
      mesh_query<MESHK>(o, d, t, u, v, id);
      S.mesh_res[4 * c + 0] = t; S.mesh_res[4 * c + 1] = u; S.mesh_res[4 * c + 2] = v; S.mesh_res[4 * c + 3] = __int_as_float(id);
    }
  }
}

// the mesh candidate of a ray out of mesh_res: the window walk of IntersectMeshBF (geometry.adb:276-296) over the pair
// results in triangle order (u, v of the selected triangle are re-derived when the caller needs them), or the stored
// per-ray result (LBVH scenes: c = the slot)
template <int MESHK, int NT>
ART_D void wave_mesh_result(const WaveSmem<NT, MESHK == MESHK_BVH>& S, int c, f3 o, f3 d, bool want_uv, int& id, float& t, float& u, float& v) {
  id = -1; t = 0.0f; u = 0.0f; v = 0.0f;
  if (c < 0) return;
  const bool pair_mode = (MESHK == MESHK_CONST) && (g_scn.pair_group != 0);
  if (pair_mode) {
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

// rank `slot` inside the material bucket of its warp peers and file it in that bucket's queue (LBVH scenes: the counting sort
// of the surface slots happens at the moment a slot's shadow query is resolved).  All 32 lanes call; pred selects.
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

template <int RT, int MESHK, int NT>
__global__ void __launch_bounds__(NT, WAVE_TOTAL_THREADS / NT) render_wave_kernel(const __grid_constant__ WaveArgs wa) {
  constexpr int NS = WAVE_SLOTS(NT);        // path slots of this block
  constexpr bool BVH = (MESHK == MESHK_BVH);
  extern __shared__ __align__(16) unsigned char wave_smem_raw[];
  WaveSmem<NT, BVH>& S = *reinterpret_cast<WaveSmem<NT, BVH>*>(wave_smem_raw);
  const gcore_dev::RenderArgs& a = wa.r;
  const int tid = threadIdx.x;
  const int gslot0 = blockIdx.x * NT;
  const int total_slots = gridDim.x * NT;
  const unsigned int total_items = (unsigned int)wa.units * (unsigned int)a.nslots;
  // work items are fetched from the global counter in batches that shrink towards the end of the launch (guided
  // self-scheduling): 2 NS while there is plenty, down to 64, so that the blocks run dry together
  auto fetch_batch = [&](unsigned int& base, unsigned int& count) {
    const unsigned int seen = *(volatile unsigned int*)wa.next_item;
    const unsigned int rem = (seen < total_items) ? (total_items - seen) : 0u;
    unsigned int want = rem / (2u * gridDim.x);
    want = want < 64u ? 64u : (want > 2u * NS ? 2u * NS : want);
    base = atomicAdd(wa.next_item, want);
    count = want;
  };
  const int aa = a.anti_aliasing;
  const int max_depth = a.max_depth;
  const f3 bg = mk3(a.bg[0], a.bg[1], a.bg[2]);
  unsigned int nrays = 0, nmesh = 0, nbounce = 0;
  const bool cnt = wa.count_mesh != 0;

  // LBVH scenes: the traversal state of a slot's pending ray lives in shared memory (cursor, running best hit, short stack), so
  // every mesh stage advances a DENSE queue of pending rays by one slice -- full warps on the traversal code -- and hands the
  // unfinished ones to the other stage's queue.  A slot whose ray is still in flight sits the next X / Y out.
  static_assert(!BVH || NS == NT, "LBVH scenes: one slot per thread");
  BvhStats tstat; tstat.nodes = 0; tstat.tris = 0; unsigned int ntrav = 0;      // LBVH nodes visited / triangle slots tested / traversals
  int bpar = 0;                                                                 // LBVH scenes: which cnt_m set this round's Y reads
  // one slice of slot s's pending traversal; returns true when the ray is resolved (result in mesh_res[4s..])
  auto bvh_slice = [&](int s, int p) -> bool {
    if constexpr (BVH) {
      f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
      float t_clip = 1000000.0f;
      if ((p & 3) == 2) {                                                     // shadow ray of Compute_Shadow, ray_tracer.adb:109-113
        const float ht = S.ht[s];
        const f3 hpos = mk3(o.x + ht * d.x, o.y + ht * d.y, o.z + ht * d.z);
        d = mk3(S.sx[s], S.sy[s], S.sz[s]);
        float epsilon = max3(fabsf(hpos.x), fabsf(hpos.y), fabsf(hpos.z)) * 0.000000001f;
        epsilon = max2(epsilon, 1.0e-30f);
        o = hpos + d * epsilon;
        // only a hit with t < maxDist - epsilon2 can change in_shadow (ray_tracer.adb:119-128): a mesh hit beyond it either is not
        // the closest hit or fails that test itself, so the walk is clipped there (the nearest hit below it is still exact)
        t_clip = S.sh_thi[s];
      }
      BvhRayState c;
      if (p & 4) { bvh_state_begin(g_scn.bvh, t_clip, c); ntrav++; }
      else {
        c.node = S.tnode[s]; c.sp = S.tsp[s];
        c.best_t = S.mesh_res[4 * s + 0]; c.best_u = S.mesh_res[4 * s + 1]; c.best_v = S.mesh_res[4 * s + 2]; c.best_prim = __float_as_int(S.mesh_res[4 * s + 3]);
      }
      const bool done = bvh_state_advance(g_scn.bvh, o, d, mk3(S.ix[s], S.iy[s], S.iz[s]), 0.0f, WAVE_BVH_SLICE, c, &S.tstack[s], NT,
                                          wa.tstack_ovf + (size_t)(gslot0 + s) * LBVH_OVERFLOW, tstat);
      S.mesh_res[4 * s + 0] = c.best_t; S.mesh_res[4 * s + 1] = c.best_u; S.mesh_res[4 * s + 2] = c.best_v;
      S.mesh_res[4 * s + 3] = __int_as_float(c.best_prim);
      if (done) S.pend[s] = 0;
      else { S.pend[s] = (short)(p & 3); S.tnode[s] = c.node; S.tsp[s] = (short)c.sp; }
      return done;
    } else { (void)s; (void)p; return true; }
  };
  // LBVH mesh stage: advance the rays of queue `qi` by one slice; unfinished -> the other queue, resolved shadow rays make their
  // slot ready for the Y that reads cnt_m[ypar]
  auto bvh_stage = [&](int qi, int ypar) {
    if constexpr (BVH) {
      const int n = S.cnt_t[qi];
      for (int j0 = 0; j0 < n; j0 += NT) {
        const int j = j0 + tid;
        bool more = false, ready = false; int s = 0, bucket = 0;
        if (j < n) {
          s = S.q_t[qi][j];
          const int p = S.pend[s];
          const bool done = bvh_slice(s, p);
          more = !done;
          ready = done && ((p & 3) == 2);
          bucket = (S.state[s] >> ST_BUCKET_SHIFT) & 7;
        }
        queue_push(more, S.q_t[qi ^ 1], &S.cnt_t[qi ^ 1], s);
        bucket_push(ready, bucket, s, S.cnt_m[ypar], &S.q_sb[0][0], NT);
      }
    } else { (void)qi; (void)ypar; }
  };

  for (int s = tid; s < NS; s += NT) {
    S.item[s] = -1; S.state[s] = 0; S.q_r[0][s] = (short)s;
    if constexpr (BVH) S.pend[s] = 0; else S.cand[s] = -1;
  }
  if (tid == 0) {
    S.cnt_r[0] = NS; S.cnt_r[1] = 0; S.retired = 0; S.exhausted = 0;
    if constexpr (BVH) { S.cnt_t[0] = 0; S.cnt_t[1] = 0; for (int b = 0; b < 5; b++) { S.cnt_m[0][b] = 0; S.cnt_m[1][b] = 0; } }
    else { S.cnt_cs = 0; S.cnt_cp = 0; for (int b = 0; b < 5; b++) S.cnt_m[b] = 0; }
    unsigned int base, count;
    fetch_batch(base, count);
    if (base >= total_items) { S.exhausted = 1; S.pool_next = 0; S.pool_end = 0; }
    else { S.pool_next = (int)base; S.pool_end = (int)min(base + count, total_items); }
  }
  __syncthreads();

  int par = 0;
  for (;;) {
    // ================================================================= X: closest hit, path end or light sample
    if (tid == 0) {
      if constexpr (BVH) S.cnt_t[1] = 0;                                      // consumed by M2 of the previous round
      else S.cnt_cp = 0;
      if (!S.exhausted && S.pool_next >= S.pool_end) {                        // refill the item pool (fetched from in Y only)
        unsigned int base, count;
        fetch_batch(base, count);
        if (base >= total_items) S.exhausted = 1;
        else { S.pool_next = (int)base; S.pool_end = (int)min(base + count, total_items); }
      }
    }
    for (int s = tid; s < NS; s += NT) {
      int st = S.state[s];
      // LBVH scenes: a slot whose path ray is still being traversed, or that waits for Y (shadow ray in flight), sits this X out
      bool on;
      if constexpr (BVH) on = ((st & (ST_HASRAY | ST_SURF)) == ST_HASRAY) && S.pend[s] == 0;
      else on = (st & ST_HASRAY) != 0;
      bool ended = false, push_cs = false, surf = false;
      int bucket = -1 - (tid & 31);                                           // distinct per lane: no peers in the match below
      if (on) {
        const f3 o = mk3(S.ox[s], S.oy[s], S.oz[s]), d = mk3(S.dx[s], S.dy[s], S.dz[s]);
        const f3 inv = mk3(S.ix[s], S.iy[s], S.iz[s]);                        // 1.0/direction (geometry.adb:154-156), computed when the ray was made
        int tri_id; float tri_t, tri_u, tri_v;
        int cidx;
        if constexpr (BVH) cidx = s; else cidx = (int)S.cand[s];
        wave_mesh_result<MESHK, NT>(S, cidx, o, d, true, tri_id, tri_t, tri_u, tri_v);
        DHit h;
        closest_with_mesh<MESHK>(o, d, inv, tri_id, tri_t, tri_u, tri_v, true, h);
        nrays++;
        st &= ~(ST_SURF | (7 << ST_BUCKET_SHIFT));
        f3 Lterm = mk3(0.f, 0.f, 0.f);
        if (!h.is_hit) {                                                      // integrators.adb:98,151,218
          ended = true;
        } else if (g_scn.mats[h.mat].type == MAT_LIGHT) {
          ended = true;
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
          surf = true;
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
            push_cs = mesh_candidate<MESHK>(so, sinv, cnt, nmesh);
          }
          const int type = g_scn.mats[h.mat].type;
          bucket = (type == MAT_LAMBERT) ? 0 : (type == MAT_PHONG) ? 1 : (type == MAT_GLASS) ? 2 : (type == MAT_MIRROR) ? 3 : 4;
          st |= bucket << ST_BUCKET_SHIFT;
        }
        if (ended) {
          S.lx[s] = Lterm.x; S.ly[s] = Lterm.y; S.lz[s] = Lterm.z;
          st = (st & ~ST_HASRAY) | ST_FOLD;
        }
        S.state[s] = st;
      }
      queue_push(ended, S.q_r[par], &S.cnt_r[par], s);
      if constexpr (BVH) {
        // a surface slot either starts its shadow traversal (queue of M1) or -- no mesh candidate, or no shadow ray at all -- is ready for Y
        if (surf) { if (push_cs) S.pend[s] = 2 | 4; else S.mesh_res[4 * s + 3] = __int_as_float(-1); }
        queue_push(surf && push_cs, S.q_t[0], &S.cnt_t[0], s);
        bucket_push(surf && !push_cs, bucket, s, S.cnt_m[bpar], &S.q_sb[0][0], NT);
      } else {
        if (RT != GCORE_PT_STUPID) {
          const int c = queue_push(push_cs, S.q_cs, &S.cnt_cs, s);
          if (on) S.cand[s] = (short)c;                                       // -1 when the shadow ray skips the mesh
        }
        // counting sort by material type: rank inside the bucket
        const unsigned m = __match_any_sync(0xffffffffu, bucket);
        if (bucket >= 0) {
          const int lane = tid & 31, leader = __ffs(m) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&S.cnt_m[bucket], __popc(m));
          base = __shfl_sync(m, base, leader);
          S.mrank[s] = (short)(base + __popc(m & ((1u << lane) - 1u)));
        }
      }
    }
    __syncthreads();

    // ================================================================= M1: mesh stage of the shadow rays
    if constexpr (BVH) {
      bvh_stage(0, bpar);
    } else {
      // scatter of the material sort + (ray, triangle) pairs
      const int c0 = S.cnt_m[0], c1 = S.cnt_m[1], c2 = S.cnt_m[2], c3 = S.cnt_m[3];
      for (int s = tid; s < NS; s += NT) {
        const int st = S.state[s];
        if ((st & (ST_SURF | ST_HASRAY)) == (ST_SURF | ST_HASRAY)) {
          const int b = (st >> ST_BUCKET_SHIFT) & 7;
          const int off_b = (b == 0) ? 0 : (b == 1) ? c0 : (b == 2) ? c0 + c1 : (b == 3) ? c0 + c1 + c2 : c0 + c1 + c2 + c3;
          S.q_s[off_b + S.mrank[s]] = (short)s;
        }
      }
      if (RT != GCORE_PT_STUPID) wave_mesh_stage<MESHK, true, NT>(S, S.q_cs, S.cnt_cs, tid);
    }
    __syncthreads();

    // ================================================================= Y: shade (slots sorted by material) + regenerate (queue R)
    {
      int c0, c1, c2, c3, c4;
      if constexpr (BVH) { c0 = S.cnt_m[bpar][0]; c1 = S.cnt_m[bpar][1]; c2 = S.cnt_m[bpar][2]; c3 = S.cnt_m[bpar][3]; c4 = S.cnt_m[bpar][4]; if (tid == 0) S.cnt_t[0] = 0; }
      else { c0 = S.cnt_m[0]; c1 = S.cnt_m[1]; c2 = S.cnt_m[2]; c3 = S.cnt_m[3]; c4 = S.cnt_m[4]; if (tid == 0) S.cnt_cs = 0; }
      // surface slots sorted by material type (buckets back to back: only the warps on a bucket boundary see two types),
      // one padding gap, then queue R
      const int n_surf = c0 + c1 + c2 + c3 + c4;
      const int n_m = (n_surf + 31) & ~31;
      const int n_r = S.cnt_r[par];
      const int n_all = n_m + ((n_r + 31) & ~31);
      for (int i = tid; i < n_all; i += NT) {
        bool ended = false, push_cp = false, starved = false, retire = false, newray = false;
        int s = 0;
        if (i < n_m) {
          // ------------------------------------------------------------- shade
          const bool on = i < n_surf;
          if (on) {
            if constexpr (BVH) {
              int b = 0, off = 0;
              if (i >= c0) { b = 1; off = c0; if (i >= c0 + c1) { b = 2; off = c0 + c1; if (i >= c0 + c1 + c2) { b = 3; off = c0 + c1 + c2; if (i >= c0 + c1 + c2 + c3) { b = 4; off = c0 + c1 + c2 + c3; } } } }
              s = S.q_sb[b][i - off];
            } else s = S.q_s[i];
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
              int tri_id; float tri_t, tri_u, tri_v;
              int cidx;
              if constexpr (BVH) cidx = s; else cidx = (int)S.cand[s];
              wave_mesh_result<MESHK, NT>(S, cidx, so, sdir, false, tri_id, tri_t, tri_u, tri_v);
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
            wa.stackA[(size_t)depth * total_slots + gslot0 + s] = make_float4(explicitColor.x, explicitColor.y, explicitColor.z, W.x);
            wa.stackB[(size_t)depth * total_slots + gslot0 + s] = make_float2(W.y, W.z);
            depth++;
            const f3 off3 = (sign_ada(cosTheta) * hn) * 1.0e-5f;              // :122,189,297
            const f3 no = hpos + off3, nd = ms.direction;
            S.ox[s] = no.x; S.oy[s] = no.y; S.oz[s] = no.z; S.dx[s] = nd.x; S.dy[s] = nd.y; S.dz[s] = nd.z;
            S.prev_pdf[s] = ms.pdf;
            st = (st & ~(ST_DEPTH_MASK | ST_SPEC | ST_SURF | ST_READY)) | depth | (ms.pureSpecular ? ST_SPEC : 0);
            if (depth == max_depth) {                                         // recursion_level = 0 -> (0,0,0)
              S.lx[s] = 0.f; S.ly[s] = 0.f; S.lz[s] = 0.f;
              st = (st & ~ST_HASRAY) | ST_FOLD;
              ended = true;
            } else {
              newray = true;
              const f3 ninv = mk3(1.0f / nd.x, 1.0f / nd.y, 1.0f / nd.z);
              S.ix[s] = ninv.x; S.iy[s] = ninv.y; S.iz[s] = ninv.z;
              push_cp = mesh_candidate<MESHK>(no, ninv, cnt, nmesh);
            }
            S.state[s] = st;
          }
        } else {
          // ------------------------------------------------------------- regenerate
          const int r = i - n_m;
          const bool on = r < n_r;
          bool need = false, padded = false;
          int st = 0, item = -1;
          if (on) {
            s = S.q_r[par][r];
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
                  A[j] = wa.stackA[(size_t)k * total_slots + gslot0 + s];
                  B[j] = wa.stackB[(size_t)k * total_slots + gslot0 + s];
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
              f3 color = mk3(S.cr[s], S.cg[s], S.cb[s]);
              bool done;
              if (aa) {
                if (sub == 0) color = bg;                                     // :42
                color = color + L;                                            // :47
                done = (sub == 3);
              } else { color = L; done = true; }                              // :60
              if (done) { wa.colbuf[item] = make_float4(color.x, color.y, color.z, 0.0f); item = -1; }
              else { S.cr[s] = color.x; S.cg[s] = color.y; S.cb[s] = color.z; }
              st = (st & ~(ST_FOLD | (3 << ST_SUB_SHIFT))) | (((sub + 1) & 3) << ST_SUB_SHIFT);
            }
            need = item < 0;
            if (!need) S.sample[s] = S.sample[s] + 1u;                        // next sub-pixel sample of the same pass
          }
          {                                                                   // next work item (pass, pixel): one atomic per warp
            const unsigned nm = __ballot_sync(0xffffffffu, need);
            if (nm) {
              const int lane = tid & 31, leader = __ffs(nm) - 1;
              int base = 0;
              if (lane == leader) base = atomicAdd(&S.pool_next, __popc(nm));
              base = __shfl_sync(0xffffffffu, base, leader);
              if (need) {
                const int got = base + __popc(nm & ((1u << lane) - 1u));
                starved = true;
                if (got < S.pool_end) {
                  const int q = got % a.nslots, unit = got / a.nslots;
                  const WavePixel px = wave_slot_to_pixel(a, q);
                  if (px.active) {                                            // (an item of a padding pixel is dropped; the slot asks again next round)
                    starved = false;
                    item = got;
                    S.pixel[s] = (unsigned int)(px.y * a.width + px.x);
                    S.sample[s] = (unsigned int)(a.first_sample + (aa ? unit * 4 : unit));
                  } else padded = true;
                }
                st &= ~(3 << ST_SUB_SHIFT);
              }
            }
          }
          if (on) {
            S.item[s] = item;
            if (!starved) {
              const int sub = (st >> ST_SUB_SHIFT) & 3;
              const unsigned int pixel = S.pixel[s];
              const int py = (int)(pixel / (unsigned int)a.width), pxx = (int)(pixel - (unsigned int)py * (unsigned int)a.width);
              const f3 o = ld3(g_scn.cam_pos);
              const f3 d = wave_eye_ray_dir(a, pxx, py, aa ? sub : -1);
              S.ox[s] = o.x; S.oy[s] = o.y; S.oz[s] = o.z; S.dx[s] = d.x; S.dy[s] = d.y; S.dz[s] = d.z;
              S.prev_pdf[s] = 1.0f;                                           // StartSample, materials.ads:25
              st = (st & (3 << ST_SUB_SHIFT)) | ST_SPEC | ST_HASRAY;          // depth 0
              newray = true;
              const f3 cinv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
              S.ix[s] = cinv.x; S.iy[s] = cinv.y; S.iz[s] = cinv.z;
              push_cp = mesh_candidate<MESHK>(o, cinv, cnt, nmesh);
            } else {
              st &= ~(ST_HASRAY | ST_FOLD);
              retire = (S.exhausted != 0) && !padded;                         // no work left anywhere: the slot retires
            }
            S.state[s] = st;
          }
        }
        if constexpr (BVH) {
          if (newray) { if (push_cp) S.pend[s] = 1 | 4; else S.mesh_res[4 * s + 3] = __int_as_float(-1); }
          queue_push(newray && push_cp, S.q_t[1], &S.cnt_t[1], s);            // the new path ray starts in M2
        } else {
          const int c = queue_push(push_cp, S.q_cp, &S.cnt_cp, s);
          if (newray) S.cand[s] = (short)c;
        }
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
    if constexpr (BVH) {
      if (tid < 5) S.cnt_m[bpar][tid] = 0;                                    // consumed by Y; shadow rays resolved from here on count for the next Y
      bvh_stage(1, bpar ^ 1);
      bpar ^= 1;
    } else {
      if (tid < 5) S.cnt_m[tid] = 0;
      wave_mesh_stage<MESHK, false, NT>(S, S.q_cp, S.cnt_cp, tid);
    }
    par ^= 1;
    __syncthreads();
  }

  // statistics: one atomic per warp and counter
  unsigned int total_rays = __reduce_add_sync(0xffffffffu, nrays);
  unsigned int total_mesh = __reduce_add_sync(0xffffffffu, nmesh);
  unsigned int total_bnc  = __reduce_add_sync(0xffffffffu, nbounce);
  if ((threadIdx.x & 31) == 0 && total_rays) {
    atomicAdd(&a.counters[0], (unsigned long long)total_rays);
    atomicAdd(&a.counters[1], (unsigned long long)total_mesh);
    atomicAdd(&a.counters[2], (unsigned long long)total_bnc);
  }
  if (BVH) {
    unsigned int tn = __reduce_add_sync(0xffffffffu, tstat.nodes), tt = __reduce_add_sync(0xffffffffu, tstat.tris), tv = __reduce_add_sync(0xffffffffu, ntrav);
    if ((threadIdx.x & 31) == 0 && tn) {
      atomicAdd(&a.counters[3], (unsigned long long)tn); atomicAdd(&a.counters[4], (unsigned long long)tt); atomicAdd(&a.counters[5], (unsigned long long)tv);
    }
  }
}

// K6a: acc[q] := colour(pass k, q) + acc[q] for k = 0 .. units-1 -- the additions of integrators.adb:51/:63 in pass order
__global__ void __launch_bounds__(256) fold_kernel(const __grid_constant__ gcore_dev::RenderArgs a, const float4* __restrict__ colbuf, int units) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= a.nslots) return;
  const WavePixel px = wave_slot_to_pixel(a, q);
  if (!px.active) return;
  float4 acc4 = a.accum[q];
  f3 acc = mk3(acc4.x, acc4.y, acc4.z);
#pragma unroll 4
  for (int k = 0; k < units; k++) {
    const float4 c = __ldcs(&colbuf[(size_t)k * a.nslots + q]);
    acc = mk3(c.x, c.y, c.z) + acc;
  }
  a.accum[q] = make_float4(acc.x, acc.y, acc.z, acc4.w);
}

}  // namespace art
