/*
 * gcore.h -- C ABI of libgcore_b200.so, the B200-native render core for ada-ray-tracer.
 *
 * The first six functions are the reference's geometry-core seam, byte-compatible with
 * cpp/embree_connect.cpp (imported by scene_hydra_embree.adb:37-66 with pragma Import (C, ...)); the rest
 * extends the seam so that a whole render pass (integrator loop + scene queries + shading + accumulation,
 * ray_tracer-integrators.adb:25-301) runs on the GPU behind one call instead of one FFI crossing per ray.
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer it passes and the library keeps
 * nothing from it after the call returns.  Functions returning int return 0 on success and a negative
 * GCORE_E_* code on failure; gcore_last_error() gives the text.  Nothing throws or exits across the ABI.
 * There is no CPU fallback: without a CUDA device every compute call fails with GCORE_E_CUDA.
 */
#ifndef GCORE_B200_H
#define GCORE_B200_H

#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCORE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * 1. Legacy geometry-core seam (drop-in for cpp/embree_connect.cpp)
 * ---------------------------------------------------------------------------------------------- */

/* replaces embree_connect.cpp:186-194 (struct HitCpp, 36 bytes; written only on a hit) */
typedef struct gcore_HitCpp {
  int   primIndex;
  int   geomIndex;
  int   instIndex;
  float t;
  float normal[3];     /* unnormalised geometric normal cross(B-A, C-A), like Embree's Ng          */
  float texCoord[2];   /* barycentrics of vertex B and vertex C, like Embree's (u, v)             */
} gcore_HitCpp;

/* replaces embree_connect.cpp:60-67 : drop every mesh/instance and start a new scene */
GCORE_API void gcore_init_and_clear(void);
/* replaces embree_connect.cpp:51-58 */
GCORE_API void gcore_destroy(void);
/* replaces embree_connect.cpp:69-144 : returns the mesh id 0,1,2,... (-1 on failure; the reference returns an
 * ambiguous 0).  Triangles with an index outside [0, vertexNum) are rewritten to (vertexNum-1) x3 as the
 * reference does (:98-109).  vertexNum counts float3 vertices, indicesNum counts ints (3 per triangle). */
GCORE_API int  gcore_add_mesh_3f(const float* a_vertices3f, int a_vertexNum, const int* a_indices, int a_indicesNum);
/* replaces embree_connect.cpp:147-184 : 16 floats per matrix, the first 12 are a 3x4 row-major transform */
GCORE_API void gcore_instance_meshes(int a_geomId, const float* a_matrices16f, int a_matrixNum);
/* replaces embree_connect.cpp:241-244 : flattens the instances and builds the LBVH on the GPU */
GCORE_API void gcore_commit_scene(void);
/* replaces embree_connect.cpp:196-238 : nearest hit with t in (t_near, t_far); re-entrant (callable from any number
 * of host threads, scene_hydra_embree.adb:426-446 does so from every Ada task).  One ray per call costs a kernel
 * launch; use gcore_bvh_closest_hit_batch for throughput.
 * Semantics of the triangle test: by default the REFERENCE's one-sided Moeller-Trumbore test (geometry.adb:231-263, back
 * faces are not hit) so that the seam agrees with the analytic build; gcore_set_seam_two_sided(1) selects Embree's
 * two-sided rtcIntersect1 behaviour.  With no gcore_instance_meshes call before the commit every mesh is instanced once
 * with the identity (Embree would trace an empty top-level scene, cpp/embree_connect.cpp:147-184). */
GCORE_API bool gcore_closest_hit(const float a_rayPos[3], const float a_rayDir[3], float t_near, float t_far,
                                 gcore_HitCpp* pHit);

/* ---- extension of the seam ---- */
/* one-sided (0, default: reference semantics) or two-sided (1: Embree semantics) triangle test of the seam queries */
GCORE_API int gcore_set_seam_two_sided(int enable);
/* material id (index into the gcore_set_materials table, default 2 like IntersectMeshBF geometry.adb:312) and optional
 * per-vertex normals of a mesh added with gcore_add_mesh_3f; call before gcore_commit_scene.  The reference reads the
 * materials library but returns null for every material (scene_hydra_embree.adb:192-202) and hands the meshes over
 * without normals (:253-268); without normals the normalised geometric normal cross(B-A, C-A) is shaded. */
GCORE_API int gcore_set_mesh_material(int meshId, int matId);
GCORE_API int gcore_set_mesh_normals_3f(int meshId, const float* a_normals3f, int a_vertexNum);
/* 1: gcore_render* / gcore_closest_hit_batch / gcore_primary_hits trace the COMMITTED scene (all meshes and instances, per-mesh
 * materials, nearest-hit semantics) together with the analytic primitives of section 2; 0 (default): the gcore_set_mesh mesh.
 * prim_index of a triangle hit is then the index into the flattened scene (gcore_get_committed_scene). */
GCORE_API int gcore_use_committed_scene(int enable);
/* the committed scene as the devices see it: world-space vertices, int3 triangles, 9 normal floats and one material id per
 * flattened triangle (instances in declaration order, each mesh's triangles in order).  Any pointer may be NULL. */
GCORE_API int gcore_get_committed_scene(float* pos3, int32_t* tris3, float* normals9, int32_t* mats, int* nverts, int* ntris);

/* ------------------------------------------------------------------------------------------------
 * 2. Scene description of the analytic build (replaces the state held by scene.ads:61-80 Render_Scene)
 * ---------------------------------------------------------------------------------------------- */

typedef struct gcore_sphere { float pos[3]; float r; int32_t mat; } gcore_sphere;            /* geometry.ads:21-25 */
typedef struct gcore_cornell {                                                                /* geometry.ads:46-53 */
  float bmin[3], bmax[3]; int32_t mat_indices[6]; float normals[18];
} gcore_cornell;
typedef struct gcore_flatlight {                                                              /* geometry.ads:38-44 */
  float boxMin[3], boxMax[3], normal[3], intensity[3]; float surfaceArea;
} gcore_flatlight;

enum { GCORE_MAT_NULL = 0, GCORE_MAT_LIGHT = 1, GCORE_MAT_LAMBERT = 2, GCORE_MAT_MIRROR = 3,
       GCORE_MAT_GLASS = 4, GCORE_MAT_PHONG = 5 };
typedef struct gcore_material {                                                               /* materials.ads:58-130 */
  int32_t type;
  float   color[3];     /* kd (Lambert) | reflection (mirror, glass, Phong) */
  float   color2[3];    /* transparency (glass)                             */
  float   param;        /* ior (glass) | cosPower (Phong)                   */
  int32_t light;        /* MaterialLight.lref: 0 = the scene light, -1 = null */
} gcore_material;

enum { GCORE_LIGHT_RECT = 0, GCORE_LIGHT_SPHERE = 1 };
typedef struct gcore_light {                                                                  /* lights.ads:36-55 */
  int32_t shape;
  float   boxMin[3], boxMax[3], normal[3];      /* AreaLight   */
  float   center[3], radius;                    /* SphereLight */
  float   intensity[3];
  float   surfaceArea;
} gcore_light;

typedef struct gcore_camera { float pos[3], lookAt[3], up[3], matrix[16]; } gcore_camera;     /* scene.ads:27-32 */

enum { GCORE_MESH_REFERENCE_BF = 0,   /* geometry.adb:266-323 semantics (bbox reject + windowed index-order loop), <= 64 triangles */
       GCORE_MESH_NEAREST      = 1 }; /* nearest front-facing hit, lowest index on ties; LBVH above 64 triangles                   */

/* replaces scene.adb:92-148,182-192 (spheres), scene.ads:75-80 (box), scene.adb:130-136 (flat light geometry) */
GCORE_API int gcore_set_analytic_scene(const gcore_sphere* spheres, int nspheres, const gcore_cornell* box,
                                       const gcore_flatlight* flatlight);
/* replaces scene.adb:153-185 */
GCORE_API int gcore_set_materials(const gcore_material* materials, int n);
/* replaces scene.adb:102-128 */
GCORE_API int gcore_set_light(const gcore_light* light);
/* replaces scene.adb:212-215 */
GCORE_API int gcore_set_camera(const gcore_camera* cam);
/* the shaded mesh (Render_Scene.mymesh): world-space positions, object-space normals (the reference never
 * rotates them, geometry.adb:605), int3 triangles.  bbox6 = min.xyz,max.xyz as folded by the loader, or NULL to
 * fold it here the way geometry.adb:593-603 does (starting from zero). */
GCORE_API int gcore_set_mesh(const float* pos3, const float* nrm3, int nverts, const int32_t* tris3, int ntris,
                             const float* bbox6, int mesh_mode);
/* replaces geometry.adb:499-609 LoadMeshFromVSGF + the transform of scene.adb:194-206: native .vsgf reader;
 * matrix16 is a row-major 4x4 applied to the positions; the result becomes the shaded mesh */
GCORE_API int gcore_set_mesh_vsgf(const char* path, const float* matrix16, int mesh_mode);
/* replaces scene.adb:24-27,89-217 Scene.Init of the default build: Cornell box, 2 spheres + sphere light,
 * materials table, camera, pyramid2.vsgf under T*R*S */
GCORE_API int gcore_scene_init_cornell(const char* vsgf_path);
/* read back what the device will see (for parity checks): any pointer may be NULL */
GCORE_API int gcore_get_mesh(float* pos3, float* nrm3, int32_t* tris3, float* bbox6, int* nverts, int* ntris);

/* ------------------------------------------------------------------------------------------------
 * 3. Scene queries in batches (parity tests, Scene.Find_Closest_Hit scene.adb:56-86)
 * ---------------------------------------------------------------------------------------------- */

typedef struct gcore_ray { float origin[3], direction[3]; } gcore_ray;                        /* geometry.ads:15-19 */
enum { GCORE_PRIM_PLANE = 0, GCORE_PRIM_SPHERE = 1, GCORE_PRIM_TRIANGLE = 2, GCORE_PRIM_QUAD = 3 };
typedef struct gcore_hit {                                                                    /* geometry.ads:57-67 */
  int32_t is_hit, prim_type, prim_index, matId, mat;
  float   t, normal[3], tx, ty;
} gcore_hit;

/* full scene query (spheres, Cornell box, flat light, mesh) for n host rays -> n host hit records */
GCORE_API int gcore_closest_hit_batch(int n, const gcore_ray* rays, gcore_hit* hits);
/* triangle-only query on the committed legacy scene (section 1) for n host rays; found[i] != 0 on a hit */
GCORE_API int gcore_bvh_closest_hit_batch(int n, const gcore_ray* rays, float t_near, float t_far,
                                          gcore_HitCpp* hits, int32_t* found);
/* gcore_closest_hit combines the calls that are pending across the host's threads into one kernel launch (flat combining):
 * out2 = {launches so far, rays answered so far}; rays / launches = the average batch */
GCORE_API int gcore_seam_stats(uint64_t* out2);
/* the batches are answered by a RESIDENT kernel that takes them from a mailbox in mapped pinned host memory (no launch, no stream
 * synchronisation per call; it leaves the GPU after 2 ms without a call and is started again on demand; GCORE_SEAM_SERVER=0 in the
 * environment goes back to one launch per batch): out2 = {times the resident kernel was started, batches it answered} */
GCORE_API int gcore_seam_server_stats(uint64_t* out2);
/* occlusion query on the committed scene for n host rays (the rtcOccluded1 counterpart, cpp/embree3/rtcore_scene.h:128, which the
 * reference's shim never exposed): occluded[i] = 1 if ANY triangle is hit with t in (t_near, t_far) */
GCORE_API int gcore_bvh_occluded_batch(int n, const gcore_ray* rays, float t_near, float t_far, int32_t* occluded);

/* ------------------------------------------------------------------------------------------------
 * 4. Rendering (replaces Ray_Tracer: ray_tracer.ads:20-48, ray_tracer.adb:240-320,
 *    ray_tracer-integrators.adb:25-301)
 * ---------------------------------------------------------------------------------------------- */

enum { GCORE_RT_DEBUG = 0, GCORE_RT_WHITTED = 1, GCORE_PT_STUPID = 2, GCORE_PT_SHADOW = 3, GCORE_PT_MIS = 4 };  /* ray_tracer.ads:40 */

typedef struct gcore_render_config {          /* the package-level variables of ray_tracer.ads:20-32 */
  int32_t  render_type;                       /* Init_Render                                      */
  int32_t  anti_aliasing;                     /* Anti_Aliasing_On: 4 fixed sub-pixel samples      */
  int32_t  max_depth;                         /* Max_Trace_Depth (<= 16)                          */
  uint32_t seed;                              /* Philox key word 0                                */
  float    background[3];                     /* Background_Color                                 */
} gcore_render_config;

typedef struct gcore_render_stats {
  uint64_t rays;            /* scene queries TRACED by the last render call (primary + bounce + shadow)  */
  uint64_t samples;         /* pixel samples added by the last render call                             */
  uint64_t mesh_queries;    /* of `rays`, those that ran the mesh stage (statistics level 2: that entered the reference's mesh bbox) */
  uint64_t bounces;         /* shaded (non-light) surface interactions of the last render call         */
  uint64_t reserved0;       /* 0 */
  float    kernel_ms;       /* device time of the render kernel(s), CUDA events                        */
  float    total_ms;        /* device time including resolve and copies                                */
  int32_t  spp;             /* samples per pixel accumulated so far                                    */
  int32_t  kernel_launches; /* kernels launched by the last call                                       */
} gcore_render_stats;

/* select the (first) CUDA device for this process (default 0) and the tile partition for multi-PROCESS rendering:
 * this process renders the 32x32 tiles t with t % world == rank.  Resets the accumulator. */
GCORE_API int gcore_set_device(int device);
GCORE_API int gcore_set_partition(int rank, int world);
/* multi-GPU inside ONE host process (the reference host is a single Ada process, ray_tracer.adb:264-277): devices
 * [device, device + n) each hold the whole scene and render the tiles t with t % n == d; every call of this header works
 * unchanged -- gcore_render_units launches on all devices, gcore_read_frame / gcore_render gather the slot-major
 * accumulators on the first device with one grouped ncclSend/ncclRecv and resolve there.  Results are bit-identical for
 * every n.  Resets the accumulator.  Exclusive with gcore_set_partition(world > 1). */
GCORE_API int gcore_set_gpu_count(int n);
GCORE_API int gcore_get_gpu_count(void);
/* device time of the render kernels of the last gcore_render_units(.., stats != NULL) per device: out[0..n) milliseconds */
GCORE_API int gcore_device_kernel_ms(float* out, int n);
/* Init_Render (ray_tracer.adb:197-200) + the config variables */
GCORE_API int gcore_init_render(const gcore_render_config* cfg);
/* Resize_Viewport (ray_tracer.adb:297-320): allocate + zero the accumulator, spp := 0 */
GCORE_API int gcore_resize_viewport(int width, int height);
/* `units` DoPass units on the device; each adds 4 (AA) or 1 spp, exactly like one Path_Trace_Thread doing one
 * pass (Render_Pass = Threads_Num units).  No host transfer. */
GCORE_API int gcore_render_units(int units, gcore_render_stats* stats);
/* counters accumulated since the last gcore_resize_viewport (rays, mesh_queries, bounces, LBVH nodes visited): lets a caller
 * enqueue many gcore_render_units(units, NULL) calls without a host sync and read the totals once.  out4 = 4 x uint64 */
GCORE_API int gcore_get_counters(uint64_t* out4);
/* LBVH work of the render kernel since the last gcore_resize_viewport: out3 = {nodes visited (64 B each), triangle slots
 * tested (48 B each), traversals started} -- the measured visit counts of the traversal roofline */
GCORE_API int gcore_get_bvh_counters(uint64_t* out3);
/* statistics level of the render kernel: 1 (default) rays / bounces are exact and mesh_queries = rays that ran the mesh
 * stage; 2 = mesh_queries counts the rays that enter the reference's mesh bounding box (geometry.adb:272-274; the 564-flop
 * queries of the roofline), which costs a bbox test per ray -- use it for an untimed counting pass */
GCORE_API int gcore_set_stats_level(int level);
/* GetSPP */
GCORE_API int gcore_get_spp(void);
/* resolve the accumulator (ray_tracer.adb:281-291) and copy to HOST buffers, row-major [y*W+x];
 * hdr4 = float4 (r,g,b,spp) per pixel, ldr = R | G<<8 | B<<16.  Either may be NULL.  Gathers first when gcore_set_gpu_count > 1;
 * not valid under a multi-process partition (world > 1: gather yourself, then gcore_resolve_gathered_device). */
GCORE_API int gcore_read_frame(float* hdr4, uint32_t* ldr);
/* checkpoint / resume (the reference only keeps the accumulator in memory and rewrites the BMP every pass, test.adb:63-69):
 * load a frame saved with gcore_read_frame (row-major float4, HOST) into the accumulator of the current viewport and
 * continue at sample index `spp`; every rank loads its own tiles of the same full frame */
GCORE_API int gcore_load_frame(const float* hdr4, int spp);
/* one-shot: resize + render `units` + read back to host buffers (the end-to-end call) */
GCORE_API int gcore_render(const gcore_render_config* cfg, int width, int height, int units,
                           float* hdr4, uint32_t* ldr, gcore_render_stats* stats);
/* primary-ray cast (Debug_Ray_Tracing ray_tracer.adb:208-238 keeps only the palette colour; this returns the
 * Hit record of the same call) into a HOST array hits[y*W+x] */
GCORE_API int gcore_primary_hits(int width, int height, gcore_hit* hits);

/* ---- device-pointer variants for callers that keep data resident (PyTorch tensors, multi-GPU gather) ---- */
/* slot-major accumulator of this process: nslots float4; layout described in DESIGN.md */
GCORE_API int gcore_accum_device_ptr(void** d_accum, int64_t* nslots);
/* resolve from this process's accumulator into DEVICE buffers (world == 1) */
GCORE_API int gcore_resolve_device(void* d_hdr4, void* d_ldr);
/* resolve from `world` gathered slot-major accumulators laid out back to back in DEVICE memory */
GCORE_API int gcore_resolve_gathered_device(const void* d_all_accum, int spp, void* d_hdr4, void* d_ldr);
/* all work of this library is enqueued on this stream (cudaStream_t as void*) */
GCORE_API int gcore_stream(void** stream);
GCORE_API int gcore_synchronize(void);

/* ------------------------------------------------------------------------------------------------
 * 5. Diagnostics
 * ---------------------------------------------------------------------------------------------- */
enum { GCORE_OK = 0, GCORE_E_CUDA = -1, GCORE_E_ARG = -2, GCORE_E_STATE = -3, GCORE_E_IO = -4, GCORE_E_LIMIT = -5 };
/* copies the last error text (NUL-terminated, truncated to n) and returns its code */
GCORE_API int gcore_last_error(char* buf, int n);
/* LBVH facts of the committed/active mesh: triangles, nodes, bytes, build milliseconds */
GCORE_API int gcore_bvh_info(int64_t* ntris, int64_t* nnodes, int64_t* bytes, float* build_ms);
/* traversal micro-benchmark (K4 roofline): n HOST rays against the legacy scene (which = 0) or the shaded mesh
 * (which = 1); one counting launch fills stats3 = {nodes visited (64 B each), triangle slots tested (48 B each), hits},
 * then `repeat` plain launches are timed with CUDA events -> ms_per_launch */
GCORE_API int gcore_bvh_bench(int n, const gcore_ray* rays, float t_near, float t_far, int which, int repeat,
                              float* ms_per_launch, uint64_t* stats3);
/* registers per thread of the render kernel for a render type (from cudaFuncGetAttributes) */
GCORE_API int gcore_kernel_regs(int render_type);

#ifdef __cplusplus
}
#endif
#endif /* GCORE_B200_H */
