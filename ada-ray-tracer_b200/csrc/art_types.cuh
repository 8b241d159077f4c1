// art_types.cuh -- plain device-side scene structures shared by the kernels and the host API.
#pragma once

#include <cuda_runtime.h>

namespace art { struct f3 { float x, y, z; }; }

#define GC_MAX_SPHERES   96
#define GC_MAX_MATERIALS 32
#define GC_MAX_BF_TRIS   64

namespace art {

enum { PRIM_PLANE = 0, PRIM_SPHERE = 1, PRIM_TRIANGLE = 2, PRIM_QUAD = 3 };
enum { MAT_NULL = 0, MAT_LIGHT = 1, MAT_LAMBERT = 2, MAT_MIRROR = 3, MAT_GLASS = 4, MAT_PHONG = 5 };
enum { LIGHT_RECT = 0, LIGHT_SPHERE = 1 };
enum { MESH_REFERENCE_BF = 0, MESH_NEAREST = 1 };

struct DMaterial { int type; float color[3]; float color2[3]; float param; int light; };

struct BvhView {                   // packed LBVH (see lbvh.cuh); all pointers are device memory
  const float4* nodes;             // 2 x float4 per internal node (32-byte quantised node, lbvh_traverse.cuh)
  const float4* tris;              // 3 x float4 per sorted triangle: (A, prim id), (e1, 0), (e2, 0)
  const float*  tri_normals;       // 9 floats per ORIGINAL triangle id: nA, nB, nC
  const int*    tri_mat;           // material id per ORIGINAL triangle id, or null: matId 2 (IntersectMeshBF, geometry.adb:312)
  int           ntris;
  int           root;              // root node index, or ~0 for a single-triangle tree
};

struct DeviceScene {
  int   nspheres, nmaterials, ntris_bf, mesh_mode;
  int   light_shape, mesh_kind;    // mesh_kind: 0 none, 1 constant-memory triangles, 2 LBVH
  int   pair_group, pad1;      // pair_group: lanes per ray in the (ray, triangle) pair stage (power of two >= ntris_bf, <= 32; 0: one ray per thread)
  float sph[GC_MAX_SPHERES][4];    // pos.xyz, r
  int   sph_mat[GC_MAX_SPHERES];
  float box_min[3], box_max[3];
  int   box_mat[6];
  float box_nrm[18];
  float fl_min[3], fl_max[3];      // FlatLight boxMin/boxMax (geometry.ads:38-44)
  float l_boxMin[3], l_boxMax[3], l_normal[3], l_center[3], l_radius, l_intensity[3], l_area;
  float cam_pos[3], cam_m[12];
  float mesh_bmin[3], mesh_bmax[3];            // the reference's bbox (folded from zeros, geometry.adb:593-603)
  float mesh_tight_min[3], mesh_tight_max[3];  // padded bounds of the vertices: a pure culling device
  DMaterial mats[GC_MAX_MATERIALS];
  float triA[GC_MAX_BF_TRIS][3], triE1[GC_MAX_BF_TRIS][3], triE2[GC_MAX_BF_TRIS][3];
  float triN[GC_MAX_BF_TRIS][9];   // nA, nB, nC
  BvhView bvh;                     // the shaded mesh (mesh_kind == 2)
  BvhView bvh_legacy;              // scene committed through the legacy gcore_* seam
};

struct DHit {                      // geometry.ads:57-67
  int   is_hit, prim_type, prim_index, matId, mat;
  float t;
  f3    normal;
  float tx, ty;
};

}  // namespace art

