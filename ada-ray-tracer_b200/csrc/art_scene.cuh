// art_scene.cuh -- device-side scene description and the scene query (Scene.Find_Closest_Hit).
//
// Replaces: scene.adb:56-86 (Find_Closest_Hit), geometry.adb:48-323 (sphere / box / Cornell / flat light /
// triangle / brute-force mesh tests).  The whole analytic scene lives in __constant__ memory (uniform
// operands, no load instructions in the inner loops); big meshes are reached through the LBVH of lbvh.cuh.
#pragma once

#include "art_math.cuh"
#include "art_types.cuh"

__constant__ art::DeviceScene g_scn;

namespace art {

// geometry.adb:146-191 IntersectBox -- returns is_hit, writes tmin/tmax; inv = (1/d.x, 1/d.y, 1/d.z) (:154-156)
ART_D bool intersect_box_inv(f3 o, f3 inv, const float* bmin, const float* bmax, float& tmin_o, float& tmax_o) {
  float inv_dir_x = inv.x;
  float inv_dir_y = inv.y;
  float inv_dir_z = inv.z;
  float lo  = (bmax[0] - o.x) * inv_dir_x;
  float hi  = (bmin[0] - o.x) * inv_dir_x;
  float lo1 = (bmax[1] - o.y) * inv_dir_y;
  float hi1 = (bmin[1] - o.y) * inv_dir_y;
  float lo2 = (bmax[2] - o.z) * inv_dir_z;
  float hi2 = (bmin[2] - o.z) * inv_dir_z;
  float tmin = min2(lo, hi);
  float tmax = max2(lo, hi);
  tmin = max2(tmin, min2(lo1, hi1));
  tmax = min2(tmax, max2(lo1, hi1));
  tmin = max2(tmin, min2(lo2, hi2));
  tmax = min2(tmax, max2(lo2, hi2));
  tmin_o = tmin; tmax_o = tmax;
  return (tmax > 0.0f) && (tmin <= tmax);
}

// geometry.adb:231-263 IntersectTriangle with edge1 = B - A, edge2 = C - A precomputed at commit time
// (same binary32 subtraction, done once).  The conjunction `v > 0 and u > 0 and u + v < 1 and t > t_min and t < t_max`
// is evaluated lazily: a lane leaves as soon as one term fails, the values that are computed are the reference's.
ART_D bool intersect_triangle(f3 o, f3 d, f3 A, f3 edge1, f3 edge2, float t_min, float t_max, float& t_o, float& u_o, float& v_o) {
  f3 pvec = cross(d, edge2);
  f3 tvec = o - A;
  float invDet = 1.0f / max2(dot(edge1, pvec), 1.0e-25f);
  float v = dot(tvec, pvec) * invDet;
  if (!(v > 0.0f)) return false;
  f3 qvec = cross(tvec, edge1);
  float u = dot(qvec, d) * invDet;
  if (!(u > 0.0f && u + v < 1.0f)) return false;
  float t = dot(edge2, qvec) * invDet;
  if (t > t_min && t < t_max) { t_o = t; u_o = u; v_o = v; return true; }
  return false;
}

// conservative slab test used ONLY to skip work whose result is known to be "no triangle hit": boxes are padded by far
// more than the rounding error of the triangle test, NaNs (0 * inf) never cull.
ART_D bool cull_box_pass(f3 o, f3 inv, const float* bmin, const float* bmax) {
  float x0 = (bmin[0] - o.x) * inv.x, x1 = (bmax[0] - o.x) * inv.x;
  float y0 = (bmin[1] - o.y) * inv.y, y1 = (bmax[1] - o.y) * inv.y;
  float z0 = (bmin[2] - o.z) * inv.z, z1 = (bmax[2] - o.z) * inv.z;
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
  return tn <= tf * 1.00001f + 1.0e-6f;
}

}  // namespace art

#include "lbvh_traverse.cuh"

namespace art {

enum { MESHK_CONST = 1, MESHK_BVH = 2 };   // where the shaded mesh lives: __constant__ triangles (<= 64) or the LBVH

// Scene.Find_Closest_Hit (scene.adb:56-86).  Candidates are evaluated in the reference's order and the nearest wins with
// strict '<'; per-candidate records are only materialised for the winner (same values), and only when the caller needs
// more than is_hit / t (shadow queries, ray_tracer.adb:117-119, read nothing else).
template <int MESHK>
ART_D void find_closest_hit(f3 o, f3 d, bool want_record, DHit& h, unsigned int& mesh_queries) {
  // --- hits(1): IntersectAllSpheres, geometry.adb:48-115
  float min_t = ART_FLT_LAST;
  int   min_i = 0;
#pragma unroll 1
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

  int   winner = 0;                       // nearestHitIndex defaults to hits'First
  float nearestHitDist = ART_FLT_LAST;
  if (sph_hit && min_t < nearestHitDist) { winner = 0; nearestHitDist = min_t; }

  // 1/direction is shared by both IntersectBox calls (geometry.adb:154-156 computes the same three quotients each time)
  const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);

  // --- hits(2): IntersectCornellBox, geometry.adb:193-229
  int box_plane = 0;
  {
    float tmin, tmax;
    if (intersect_box_inv(o, inv, g_scn.box_min, g_scn.box_max, tmin, tmax)) {
      if (tmax < nearestHitDist) {        // classification only matters if the box can still win
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

  // --- hits(3): IntersectFlatLight, geometry.adb:118-143 (only for a rect light, scene.adb:65-67)
  if (g_scn.light_shape == LIGHT_RECT) {
    float tmin = (g_scn.fl_max[1] - o.y) * inv.y;
    f3 hp = mk3(o.x + tmin * d.x, o.y + tmin * d.y, o.z + tmin * d.z);
    bool is_hit = (hp.x > g_scn.fl_min[0]) && (hp.x < g_scn.fl_max[0]) && (hp.z > g_scn.fl_min[2]) && (hp.z < g_scn.fl_max[2]) && (tmin >= 0.0f);
    if (is_hit && tmin < nearestHitDist) { winner = 2; nearestHitDist = tmin; }
  }

  // --- hits(4): mesh
  int   tri_id = 0;
  float tri_t = 0.0f, tri_u = 0.0f, tri_v = 0.0f;
  if (MESHK == MESHK_CONST) {
    if (g_scn.ntris_bf > 0) {
      const bool ref_mode = (g_scn.mesh_mode == MESH_REFERENCE_BF);
      // IntersectMeshBF, geometry.adb:266-323: bbox reject (reference mode only) ...
      float bt0, bt1;
      bool enter = ref_mode ? intersect_box_inv(o, inv, g_scn.mesh_bmin, g_scn.mesh_bmax, bt0, bt1) : true;
      if (enter) {
        mesh_queries++;
        // ... the reference's bbox also spans the origin (it is never reset, geometry.adb:593-603): skip the triangle
        // loop for rays that miss the padded tight box of the vertices -- none of the triangle tests can succeed
        if (cull_box_pass(o, inv, g_scn.mesh_tight_min, g_scn.mesh_tight_max)) {
          // every triangle in index order against the window (nearest.tmin, nearest.tmax); in reference mode a hit
          // moves the window to (t, t + 1e-6) (geometry.adb:289-293, 255-256), in nearest mode to (0, t)
          float w_min = 0.0f, w_max = 1000000.0f;
          bool any = false;
#pragma unroll 1
          for (int i = 0; i < g_scn.ntris_bf; i++) {
            float t, u, v;
            if (intersect_triangle(o, d, ld3(g_scn.triA[i]), ld3(g_scn.triE1[i]), ld3(g_scn.triE2[i]), w_min, w_max, t, u, v)) {
              if (ref_mode) { w_min = t; w_max = t + 1.0e-6f; } else { w_max = t; }
              tri_t = t; tri_u = u; tri_v = v; tri_id = i; any = true;
            }
          }
          if (any && tri_t < nearestHitDist) { winner = 3; nearestHitDist = tri_t; }
        }
      }
    }
  } else {
    int prim; float t, u, v;
    mesh_queries++;
    bvh_closest(g_scn.bvh, o, d, inv, 0.0f, 1000000.0f, prim, t, u, v);
    if (prim >= 0) {
      tri_id = prim; tri_t = t; tri_u = u; tri_v = v;
      if (tri_t < nearestHitDist) { winner = 3; nearestHitDist = tri_t; }
    }
  }

  // --- the winner's record
  h.is_hit = (winner == 0) ? (sph_hit ? 1 : 0) : 1;
  h.t = (winner == 0) ? (sph_hit ? min_t : 1.0f) : ((winner == 3) ? tri_t : nearestHitDist);
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
    // itx/ity (geometry.adb:302-303): the loader zeroes the uvs (geometry.adb:565-566) -> w*0 + v*0 + u*0
    h.tx = w * 0.0f + tri_v * 0.0f + tri_u * 0.0f;
    h.ty = h.tx;
  }
  if (h.mat < 0) h.mat = h.matId;                                               // scene.adb:80-82
}

}  // namespace art
