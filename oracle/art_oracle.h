/*
 * art_oracle.h -- C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement of the reference
 * renderer's hot path (FROL256/ada-ray-tracer, Ada sources cited per function in
 * art_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (libgcore_b200.so) never includes, links or calls anything under oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path, and neither GNAT (Ada) nor libembree3.a exists in this image, so the
 * reference itself cannot be run.  The oracle is pinned only by (i) hand-derived
 * known-answer rays (tests/test_oracle_kat.py, computed independently in float64
 * numpy from the scene constants of scene.adb), (ii) the Random123 Philox4x32-10
 * known-answer vectors, (iii) the exact bytes of data/pyramid2.vsgf.
 */
#ifndef ART_ORACLE_H
#define ART_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plain-old-data scene description (same byte layout as include/gcore.h,
 *      defined independently here) ------------------------------------------- */

typedef struct { float pos[3]; float r; int32_t mat; } ao_sphere;            /* geometry.ads:21-25  */

typedef struct {                                                             /* geometry.ads:46-53  */
  float   bmin[3], bmax[3];
  int32_t mat_indices[6];
  float   normals[18];
} ao_cornell;

typedef struct {                                                             /* geometry.ads:38-44  */
  float boxMin[3], boxMax[3], normal[3], intensity[3];
  float surfaceArea;
} ao_flatlight;

enum { AO_MAT_NULL = 0, AO_MAT_LIGHT = 1, AO_MAT_LAMBERT = 2, AO_MAT_MIRROR = 3,
       AO_MAT_GLASS = 4, AO_MAT_PHONG = 5 };

typedef struct {                                                             /* materials.ads:58-130 */
  int32_t type;
  float   color[3];    /* kd | reflection                       */
  float   color2[3];   /* transparency (glass)                  */
  float   param;       /* ior (glass) | cosPower (phong)        */
  int32_t light;       /* lref: 0 = the scene light, -1 = null  */
} ao_material;

enum { AO_LIGHT_RECT = 0, AO_LIGHT_SPHERE = 1 };

typedef struct {                                                             /* lights.ads:36-55 */
  int32_t shape;
  float   boxMin[3], boxMax[3], normal[3];   /* AreaLight   */
  float   center[3], radius;                 /* SphereLight */
  float   intensity[3];
  float   surfaceArea;
} ao_light;

typedef struct { float pos[3], lookAt[3], up[3], matrix[16]; } ao_camera;    /* scene.ads:27-32 */

enum { AO_PRIM_PLANE = 0, AO_PRIM_SPHERE = 1, AO_PRIM_TRIANGLE = 2, AO_PRIM_QUAD = 3 };

typedef struct {                                                             /* geometry.ads:57-67 */
  int32_t is_hit, prim_type, prim_index, matId, mat;
  float   t, normal[3], tx, ty;
} ao_hit;

typedef struct { float origin[3], direction[3]; } ao_ray;

enum { AO_RT_DEBUG = 0, AO_RT_WHITTED = 1, AO_PT_STUPID = 2, AO_PT_SHADOW = 3, AO_PT_MIS = 4 };
enum { AO_MESH_REFERENCE_BF = 0, AO_MESH_NEAREST = 1 };

typedef struct {                                                             /* ray_tracer.ads:20-32,40 */
  int32_t  width, height;
  int32_t  render_type;
  int32_t  anti_aliasing;
  int32_t  max_depth;
  int32_t  passes;          /* number of DoPass units; each adds 4 (AA) or 1 spp            */
  uint32_t seed;
  int32_t  first_sample;    /* spp already accumulated (sample index of the first new one)  */
  float    background[3];
  int32_t  mesh_mode;       /* AO_MESH_*                                                    */
  int32_t  threads;         /* CPU worker threads (row-striped); 0 = hardware_concurrency   */
  int32_t  y0, y1;          /* row range to render [y0,y1); 0,0 = whole image               */
} ao_params;

typedef struct ao_scene ao_scene;

ao_scene* ao_scene_new(void);
void      ao_scene_free(ao_scene*);

/* scene.adb:89-217 Init_Cornell_Box; returns 0 on success */
int  ao_scene_init_cornell(ao_scene*, const char* vsgf_path);

void ao_set_spheres  (ao_scene*, const ao_sphere*, int n);
void ao_set_cornell  (ao_scene*, const ao_cornell*);
void ao_set_flatlight(ao_scene*, const ao_flatlight*);
void ao_set_materials(ao_scene*, const ao_material*, int n);
void ao_set_light    (ao_scene*, const ao_light*);
void ao_set_camera   (ao_scene*, const ao_camera*);
/* mesh given in world space; bbox folded like geometry.adb:593-607 (starting from zero) */
void ao_set_mesh     (ao_scene*, const float* pos3, const float* nrm3, int nverts,
                      const int32_t* tris3, int ntris);
/* per-triangle material ids of the mesh set with ao_set_mesh (a scene committed through the legacy seam carries one material per
 * mesh, scene_hydra_embree.adb:192-202); without it every triangle has matId 2 (geometry.adb:312) */
void ao_set_mesh_tri_materials(ao_scene*, const int32_t* mats, int ntris);
/* geometry.adb:499-609; matrix16 row-major; returns 0 on success */
int  ao_load_vsgf    (ao_scene*, const char* path, const float* matrix16);

int  ao_get_counts   (const ao_scene*, int* nspheres, int* nmaterials, int* nverts, int* ntris);
void ao_get_spheres  (const ao_scene*, ao_sphere* out);
void ao_get_cornell  (const ao_scene*, ao_cornell* out);
void ao_get_materials(const ao_scene*, ao_material* out);
void ao_get_light    (const ao_scene*, ao_light* out);
void ao_get_camera   (const ao_scene*, ao_camera* out);
void ao_get_mesh     (const ao_scene*, float* pos3, float* nrm3, int32_t* tris3, float* bbox6);

/* scene.adb:56-86 */
void ao_closest_hit_batch(const ao_scene*, int mesh_mode, int n, const ao_ray* rays, ao_hit* hits);
/* ray_tracer.adb:61-97 + integrators.adb:38-46 ; sub = -1 centre, 0..3 AA sub-pixel */
void ao_eye_ray(const ao_scene*, int width, int height, int x, int y, int sub, ao_ray* out);
/* ray_tracer.adb:208-238 primary cast: hits[y*W+x] */
void ao_primary_hits(const ao_scene*, const ao_params*, ao_hit* hits);
/* ray_tracer.adb:100-132 */
int  ao_compute_shadow(const ao_scene*, int mesh_mode, const float hit_pos[3], const float lpos[3]);

/* integrators.adb:25-71 DoPass x passes; accum3[(y*W+x)*3+c] += ... ; returns rays traced
 * (number of Find_Closest_Hit calls) */
uint64_t ao_render(const ao_scene*, const ao_params*, float* accum3);
/* the reference's own scheme, for timing only: `tasks` threads, each one full-image pass, one global lock per pixel add
 * (ray_tracer.adb:142-194, 264-277; integrators.adb:50-52); same terms as ao_render with passes = tasks, order not deterministic */
uint64_t ao_render_task_lock(const ao_scene*, const ao_params*, int tasks, float* accum3);
/* ray_tracer.adb:281-291 + :19-57 ; ldr[y*W+x] = R | G<<8 | B<<16 */
void ao_resolve(const float* accum3, int npixels, int spp, uint32_t* ldr);
/* bitmap.adb:31-85 ; returns 0 on success */
int  ao_save_bmp(const char* path, const uint32_t* data, int width, int height);

/* deterministic elementary functions + RNG (exposed for unit tests) */
void  ao_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void  ao_sincos(float x, float* s, float* c);
float ao_pow(float x, float y);
void  ao_use_libm(int on);   /* 1: sinf/cosf/powf from libm instead of the deterministic kernels */

/* batched access to the shading pieces (tests/ada_numpy_shading.py pins them against an independent numpy restatement) */
void  ao_sincos_batch(int n, const float* x, float* s, float* c);
void  ao_expon_batch(int n, const float* left, const float* right, float* out);            /* Left ** Right */
void  ao_light_sample_batch(const ao_scene*, int n, const float* u2, const float* p3, float* out_pos3, float* out_pdf);
void  ao_light_eval_pdf_batch(const ao_scene*, int n, const float* p3, const float* dir3, const float* t, float* out);
void  ao_mat_sample_batch(const ao_scene*, int mat_id, int n, const float* r2, const float* ray_dir3, const float* normal3,
                          float* out_color3, float* out_dir3, float* out_pdf, int32_t* out_spec);
void  ao_mat_eval_batch(const ao_scene*, int mat_id, int n, const float* l3, const float* v3, const float* n3, float* out_bxdf3, float* out_pdf);

/* nearest-hit triangle query on an arbitrary mesh through the oracle's own CPU BVH
 * (semantics of SURVEY C5: nearest front-facing hit, lowest index on exact ties); binned-SAH build, single-ray traversal */
typedef struct ao_bvh ao_bvh;
ao_bvh* ao_bvh_build(const float* pos3, int nverts, const int32_t* tris3, int ntris);
void    ao_bvh_free(ao_bvh*);
/* out_prim[i] = triangle id or -1; out_t/out_u/out_v optional (may be NULL) */
void    ao_bvh_closest(const ao_bvh*, int n, const ao_ray* rays, float tmin, float tmax,
                       int32_t* out_prim, float* out_t, float* out_u, float* out_v, int threads);
void    ao_bruteforce_closest(const float* pos3, const int32_t* tris3, int ntris, int n,
                       const ao_ray* rays, float tmin, float tmax,
                       int32_t* out_prim, float* out_t, float* out_u, float* out_v);

/* the same brute-force query with Embree's two-sided triangle test (rtcIntersect1 semantics) */
void    ao_bruteforce_closest_two_sided(const float* pos3, const int32_t* tris3, int ntris, int n,
                       const ao_ray* rays, float tmin, float tmax,
                       int32_t* out_prim, float* out_t, float* out_u, float* out_v);

#ifdef __cplusplus
}
#endif
#endif
