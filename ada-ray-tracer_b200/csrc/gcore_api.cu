// gcore_api.cu -- the C ABI of libgcore_b200.so (include/gcore.h): process-wide render-core state, scene marshalling,
// kernel launches.  Replaces cpp/embree_connect.cpp (legacy seam) and the state of ray_tracer.adb / scene.adb that the
// hot path needs on the device.  No CPU compute path exists here: every query / render call launches CUDA kernels.
#include "gcore_internal.h"
#include "art_types.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

using gcore_dev::LbvhBuffers;
using gcore_dev::RenderArgs;
using gcore_dev::WaveBuffers;

namespace {

struct LegacyMesh { std::vector<float> verts; std::vector<int> idx; };
struct LegacyInstance { int mesh; float m[12]; };

struct State {
  std::recursive_mutex mu;
  int  device = 0;
  bool dev_ready = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;

  // legacy seam
  bool legacy_alive = false;
  std::vector<LegacyMesh> meshes;
  std::vector<LegacyInstance> instances;
  LbvhBuffers legacy_bvh;
  int* d_legacy_prim = nullptr;          // flattened triangle -> triangle index inside its mesh
  int* d_legacy_inst = nullptr;          // flattened triangle -> instance id
  std::vector<int> h_legacy_prim, h_legacy_inst;
  float* d_legacy_verts = nullptr; int* d_legacy_tris = nullptr;

  // analytic scene (host mirror of the __constant__ block)
  art::DeviceScene hs;
  bool scene_dirty = true;
  std::vector<float> mesh_pos, mesh_nrm; std::vector<int> mesh_tris; float mesh_bbox[6] = {0, 0, 0, 0, 0, 0};
  LbvhBuffers mesh_bvh;

  // render state
  gcore_render_config cfg{GCORE_PT_MIS, 1, 8, 0x5EED0003u, {0.f, 0.f, 0.f}};
  int width = 0, height = 0, spp = 0;
  bool palette_frame = false;                // the last pass was Debug_Ray_Tracing: resolve without 1/spp (ray_tracer.adb:245-257)
  int rank = 0, world = 1;
  float4* d_accum = nullptr; int nslots = 0;
  unsigned long long* d_rays = nullptr;      // 3 cumulative counters: rays, mesh-bbox queries, bounces
  unsigned long long last_counters[4] = {0, 0, 0, 0};
  float4* d_hdr = nullptr; uint32_t* d_ldr = nullptr; size_t out_pixels = 0;
  gcore_ray* d_batch_rays = nullptr; gcore_HitCpp* d_batch_hits = nullptr; int* d_batch_found = nullptr; size_t batch_cap = 0;
  WaveBuffers wave;                          // work buffers of the wavefront render kernel
  int stats_level = 1;
  int sm_count = 0;

  int  err_code = 0;
  std::string err_text;
};

State& S() { static State s; return s; }

int fail(int code, const std::string& text) {
  State& s = S();
  s.err_code = code; s.err_text = text;
  return code;
}
int fail_cuda(cudaError_t e, const char* where) {
  cudaGetLastError();
  return fail(GCORE_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(x, where) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, where); } while (0)

int ensure_device() {
  State& s = S();
  if (s.dev_ready) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(GCORE_E_CUDA, "no CUDA device available (libgcore_b200 has no CPU fallback)"); }
  if (s.device >= n) return fail(GCORE_E_ARG, "device index out of range");
  CU(cudaSetDevice(s.device), "cudaSetDevice");
  CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
  CU(cudaEventCreate(&s.ev0), "cudaEventCreate"); CU(cudaEventCreate(&s.ev1), "cudaEventCreate"); CU(cudaEventCreate(&s.ev2), "cudaEventCreate");
  CU(cudaMalloc(&s.d_rays, 6 * sizeof(unsigned long long)), "cudaMalloc");
  CU(cudaMemset(s.d_rays, 0, 6 * sizeof(unsigned long long)), "cudaMemset");
  s.dev_ready = true;
  s.scene_dirty = true;
  return 0;
}

void init_host_scene(art::DeviceScene& hs) {
  memset(&hs, 0, sizeof(hs));
  hs.light_shape = art::LIGHT_SPHERE;
  hs.cam_m[0] = hs.cam_m[5] = hs.cam_m[10] = 1.0f;
}

struct Init { Init() { init_host_scene(S().hs); } } g_init;

int sync_scene() {
  State& s = S();
  if (int rc = ensure_device()) return rc;
  if (s.scene_dirty) {
    s.hs.bvh.nodes = s.mesh_bvh.nodes; s.hs.bvh.tris = s.mesh_bvh.tris; s.hs.bvh.tri_normals = s.mesh_bvh.tri_normals;
    s.hs.bvh.ntris = s.mesh_bvh.ntris; s.hs.bvh.root = s.mesh_bvh.root;
    s.hs.bvh_legacy.nodes = s.legacy_bvh.nodes; s.hs.bvh_legacy.tris = s.legacy_bvh.tris; s.hs.bvh_legacy.tri_normals = nullptr;
    s.hs.bvh_legacy.ntris = s.legacy_bvh.ntris; s.hs.bvh_legacy.root = s.legacy_bvh.root;
    CU(gcore_dev::upload_scene(&s.hs, s.stream), "upload_scene");
    CU(cudaStreamSynchronize(s.stream), "upload_scene sync");     // hs may change right after we return
    s.scene_dirty = false;
  }
  return 0;
}

void free_accum() {
  State& s = S();
  if (s.d_accum) { cudaFree(s.d_accum); s.d_accum = nullptr; }
  s.nslots = 0;
}

int alloc_accum() {
  State& s = S();
  if (int rc = ensure_device()) return rc;
  if (s.width <= 0 || s.height <= 0) { free_accum(); return 0; }
  int tiles_x = (s.width + GC_TILE - 1) / GC_TILE, tiles_y = (s.height + GC_TILE - 1) / GC_TILE;
  int ntiles = tiles_x * tiles_y;
  int local_tiles = (ntiles + s.world - 1) / s.world;           // equal share per rank (padded) so gathers are uniform
  const int nslots = local_tiles * GC_TILE * GC_TILE;
  if (!s.d_accum || nslots != s.nslots) {                       // same size: keep the allocation, just clear it
    free_accum();
    CU(cudaMalloc(&s.d_accum, sizeof(float4) * (size_t)nslots), "cudaMalloc accum");
    s.nslots = nslots;
  }
  CU(cudaMemsetAsync(s.d_accum, 0, sizeof(float4) * (size_t)s.nslots, s.stream), "memset accum");
  CU(cudaMemsetAsync(s.d_rays, 0, 6 * sizeof(unsigned long long), s.stream), "memset counters");
  s.last_counters[0] = s.last_counters[1] = s.last_counters[2] = s.last_counters[3] = 0;
  s.spp = 0; s.palette_frame = false;
  return 0;
}

// work buffers of render_wave_kernel: per-slot recursion stacks (fixed size) and the (pass, pixel) colour buffer
#define GC_COLBUF_CAP_BYTES ((size_t)2 << 30)
int ensure_wave_buffers(size_t items) {
  State& s = S();
  if (!s.wave.sm_count) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, s.device), "cudaGetDeviceProperties");
    s.sm_count = prop.multiProcessorCount;
    s.wave.sm_count = s.sm_count;
    const size_t total_slots = (size_t)gcore_dev::wave_total_slots(s.sm_count);
    CU(cudaMalloc(&s.wave.stackA, sizeof(float4) * total_slots * gcore_dev::wave_max_depth()), "cudaMalloc stackA");
    CU(cudaMalloc(&s.wave.stackB, sizeof(float2) * total_slots * gcore_dev::wave_max_depth()), "cudaMalloc stackB");
    CU(cudaMalloc(&s.wave.next_item, sizeof(unsigned int)), "cudaMalloc next_item");
  }
  if (items > s.wave.colbuf_items) {
    if (s.wave.colbuf) { cudaFree(s.wave.colbuf); s.wave.colbuf = nullptr; s.wave.colbuf_items = 0; }
    CU(cudaMalloc(&s.wave.colbuf, sizeof(float4) * items), "cudaMalloc colbuf");
    s.wave.colbuf_items = items;
  }
  return 0;
}

RenderArgs make_args() {
  State& s = S();
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.width = s.width; a.height = s.height;
  a.tiles_x = (s.width + GC_TILE - 1) / GC_TILE;
  a.ntiles = a.tiles_x * ((s.height + GC_TILE - 1) / GC_TILE);
  a.rank = s.rank; a.world = s.world; a.nslots = s.nslots;
  a.anti_aliasing = s.cfg.anti_aliasing ? 1 : 0; a.max_depth = s.cfg.max_depth; a.passes = 0;
  a.seed = s.cfg.seed; a.first_sample = s.spp;
  a.bg[0] = s.cfg.background[0]; a.bg[1] = s.cfg.background[1]; a.bg[2] = s.cfg.background[2];
  const float fov = 1.57079632679489661923f;                      // Pi / 2.0            (ray_tracer.adb:63)
  a.zdir = -(float)s.width / tanf(fov / 2.0f);                     // safe_tan(fov / 2.0) (ray_tracer.adb:67, vector_math.adb:14-22)
  a.half_w = (float)s.width / 2.0f; a.half_h = (float)s.height / 2.0f;
  a.accum = s.d_accum; a.counters = s.d_rays;
  return a;
}

int ensure_out_buffers(bool hdr, bool ldr) {
  State& s = S();
  size_t px = (size_t)s.width * s.height;
  if (px != s.out_pixels) {
    if (s.d_hdr) { cudaFree(s.d_hdr); s.d_hdr = nullptr; }
    if (s.d_ldr) { cudaFree(s.d_ldr); s.d_ldr = nullptr; }
    s.out_pixels = px;
  }
  if (hdr && !s.d_hdr) CU(cudaMalloc(&s.d_hdr, sizeof(float4) * px), "cudaMalloc hdr");
  if (ldr && !s.d_ldr) CU(cudaMalloc(&s.d_ldr, sizeof(uint32_t) * px), "cudaMalloc ldr");
  return 0;
}

// ---- host-side restatement of the small pieces of scene set-up the path depends on -------------------------------
inline float min2(float a, float b) { return a < b ? a : b; }       // generic_vector_math.adb:19-26
inline float max2(float a, float b) { return a >= b ? a : b; }      // :28-35

void fold_bbox(const std::vector<float>& pos, float bbox[6]) {       // geometry.adb:593-603 (never reset: starts from the record's zeros)
  for (size_t i = 0; i + 2 < pos.size(); i += 3) {
    bbox[0] = min2(bbox[0], pos[i]); bbox[1] = min2(bbox[1], pos[i + 1]); bbox[2] = min2(bbox[2], pos[i + 2]);
    bbox[3] = max2(bbox[3], pos[i]); bbox[4] = max2(bbox[4], pos[i + 1]); bbox[5] = max2(bbox[5], pos[i + 2]);
  }
}

struct M4 { float m[4][4]; };
M4 ident() { M4 r; memset(&r, 0, sizeof(r)); for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f; return r; }
M4 mul(const M4& a, const M4& b) {                                   // generic_vector_math.adb:233-256
  M4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
    r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
  return r;
}
M4 rotation(float angle, float vx, float vy, float vz) {             // vector_math.adb:85-111
  M4 M = ident();
  float l_inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
  vx = l_inv * vx; vy = l_inv * vy; vz = l_inv * vz;
  float cos_t = cosf(angle), sin_t = sinf(angle);
  M.m[0][0] = (1.0f - cos_t) * vx * vx + cos_t;
  M.m[0][1] = (1.0f - cos_t) * vx * vy - sin_t * vz;
  M.m[0][2] = (1.0f - cos_t) * vx * vz + sin_t * vy;
  M.m[1][0] = (1.0f - cos_t) * vy * vx + sin_t * vz;
  M.m[1][1] = (1.0f - cos_t) * vy * vy + cos_t;
  M.m[1][2] = (1.0f - cos_t) * vy * vz - sin_t * vx;
  M.m[2][0] = (1.0f - cos_t) * vx * vz - sin_t * vy;
  M.m[2][1] = (1.0f - cos_t) * vz * vy + sin_t * vx;
  M.m[2][2] = (1.0f - cos_t) * vz * vz + cos_t;
  return M;
}

int install_mesh(int mesh_mode) {
  State& s = S();
  const int ntris = (int)s.mesh_tris.size() / 3;
  const int nverts = (int)s.mesh_pos.size() / 3;
  for (int v : s.mesh_tris) if (v < 0 || v >= nverts) return fail(GCORE_E_ARG, "gcore_set_mesh: triangle index out of range");
  gcore_dev::lbvh_free(&s.mesh_bvh);
  s.hs.mesh_mode = mesh_mode;
  s.hs.ntris_bf = 0; s.hs.mesh_kind = 0; s.hs.pair_group = 0;
  for (int k = 0; k < 3; k++) { s.hs.mesh_bmin[k] = s.mesh_bbox[k]; s.hs.mesh_bmax[k] = s.mesh_bbox[3 + k]; }
  s.scene_dirty = true;
  if (ntris == 0) return 0;
  {  // padded tight bounds of the vertices (culling only; 1e-4 relative + 1e-5 absolute >> rounding error of the triangle test)
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (size_t i = 0; i + 2 < s.mesh_pos.size(); i += 3)
      for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], s.mesh_pos[i + k]); mx[k] = fmaxf(mx[k], s.mesh_pos[i + k]); }
    for (int k = 0; k < 3; k++) {
      float pad = fmaxf(fabsf(mn[k]), fabsf(mx[k])) * 1.0e-4f + 1.0e-5f;
      s.hs.mesh_tight_min[k] = mn[k] - pad; s.hs.mesh_tight_max[k] = mx[k] + pad;
    }
  }
  auto P = [&](int v, int k) { return s.mesh_pos[3 * (size_t)v + k]; };
  auto N = [&](int v, int k) { return s.mesh_nrm[3 * (size_t)v + k]; };
  if (ntris <= GC_MAX_BF_TRIS) {
    for (int t = 0; t < ntris; t++) {
      int ia = s.mesh_tris[3 * t], ib = s.mesh_tris[3 * t + 1], ic = s.mesh_tris[3 * t + 2];
      for (int k = 0; k < 3; k++) {
        s.hs.triA[t][k]  = P(ia, k);
        s.hs.triE1[t][k] = P(ib, k) - P(ia, k);                     // edge1 := B - A (geometry.adb:238)
        s.hs.triE2[t][k] = P(ic, k) - P(ia, k);                     // edge2 := C - A (geometry.adb:239)
        s.hs.triN[t][k] = N(ia, k); s.hs.triN[t][3 + k] = N(ib, k); s.hs.triN[t][6 + k] = N(ic, k);
      }
    }
    s.hs.ntris_bf = ntris; s.hs.mesh_kind = 1;
    s.hs.pair_group = 0;
    if (ntris <= 8) s.hs.pair_group = ntris;            // WAVE_PAIR_TRIS of render_wave.cuh
    return 0;
  }
  if (mesh_mode == GCORE_MESH_REFERENCE_BF)
    return fail(GCORE_E_LIMIT, "GCORE_MESH_REFERENCE_BF (brute-force reference semantics) is limited to 64 triangles; use GCORE_MESH_NEAREST");
  if (int rc = ensure_device()) return rc;
  std::vector<float> n9(9 * (size_t)ntris);
  for (int t = 0; t < ntris; t++)
    for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) n9[9 * (size_t)t + 3 * c + k] = N(s.mesh_tris[3 * (size_t)t + c], k);
  float *d_v = nullptr, *d_n = nullptr; int* d_t = nullptr;
  auto cleanup = [&]() { cudaFree(d_v); cudaFree(d_n); cudaFree(d_t); };
  cudaError_t e;
  if ((e = cudaMalloc(&d_v, s.mesh_pos.size() * 4)) != cudaSuccess || (e = cudaMalloc(&d_t, s.mesh_tris.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&d_n, n9.size() * 4)) != cudaSuccess) { cleanup(); return fail_cuda(e, "cudaMalloc mesh"); }
  cudaMemcpyAsync(d_v, s.mesh_pos.data(), s.mesh_pos.size() * 4, cudaMemcpyHostToDevice, s.stream);
  cudaMemcpyAsync(d_t, s.mesh_tris.data(), s.mesh_tris.size() * 4, cudaMemcpyHostToDevice, s.stream);
  cudaMemcpyAsync(d_n, n9.data(), n9.size() * 4, cudaMemcpyHostToDevice, s.stream);
  e = gcore_dev::lbvh_build(d_v, nverts, d_t, ntris, d_n, &s.mesh_bvh, s.stream);
  cudaStreamSynchronize(s.stream);
  cleanup();
  if (e != cudaSuccess) return fail_cuda(e, "lbvh_build");
  s.hs.mesh_kind = 2;
  return 0;
}

void free_legacy_device() {
  State& s = S();
  gcore_dev::lbvh_free(&s.legacy_bvh);
  if (s.d_legacy_prim) { cudaFree(s.d_legacy_prim); s.d_legacy_prim = nullptr; }
  if (s.d_legacy_inst) { cudaFree(s.d_legacy_inst); s.d_legacy_inst = nullptr; }
  if (s.d_legacy_verts) { cudaFree(s.d_legacy_verts); s.d_legacy_verts = nullptr; }
  if (s.d_legacy_tris) { cudaFree(s.d_legacy_tris); s.d_legacy_tris = nullptr; }
  s.h_legacy_prim.clear(); s.h_legacy_inst.clear();
  s.scene_dirty = true;
}

int render_units_locked(int units, gcore_render_stats* stats) {
  State& s = S();
  if (units < 0) return fail(GCORE_E_ARG, "units < 0");
  if (!s.d_accum) return fail(GCORE_E_STATE, "gcore_resize_viewport has not been called");
  if (s.cfg.max_depth < 1 || s.cfg.max_depth > 16) return fail(GCORE_E_ARG, "max_depth must be in 1..16");
  if (int rc = sync_scene()) return rc;
  RenderArgs a = make_args();
  a.passes = units;
  CU(cudaEventRecord(s.ev0, s.stream), "event");
  int launches = 0;
  uint64_t samples = 0;
  const uint64_t pixels = (uint64_t)s.width * s.height;
  if (s.cfg.render_type == GCORE_RT_DEBUG || s.cfg.render_type == GCORE_RT_WHITTED) {
    CU(gcore_dev::launch_debug(s.hs.mesh_kind, a, s.stream), "debug_kernel");           // Debug_Ray_Tracing overwrites the buffer, spp stays
    launches = 1; samples = pixels;
    s.palette_frame = true;                                              // g_spp is not touched (ray_tracer.adb:245-260)
  } else if (units > 0) {
    const int per = s.cfg.anti_aliasing ? 4 : 1;
    // the colour buffer holds one float4 per (pass, pixel): long renders go in chunks of passes
    s.palette_frame = false;
    size_t cap_bytes = GC_COLBUF_CAP_BYTES;
    if (const char* e = getenv("GCORE_COLBUF_CAP_KB")) { long v = atol(e); if (v > 0) cap_bytes = (size_t)v << 10; }   // tests: force the chunked path on small frames
    size_t chunk = cap_bytes / (sizeof(float4) * (size_t)s.nslots);
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)units) chunk = (size_t)units;
    while ((chunk * (size_t)s.nslots) >> 31) chunk >>= 1;
    if (chunk < 1) return fail(GCORE_E_LIMIT, "viewport too large for the work-item counter");
    if (int rc = ensure_wave_buffers(chunk * (size_t)s.nslots)) return rc;
    for (int done = 0; done < units; done += (int)chunk) {
      const int n = (units - done < (int)chunk) ? (units - done) : (int)chunk;
      a.passes = n; a.first_sample = s.spp;
      CU(gcore_dev::launch_render_wave(s.cfg.render_type, s.hs.mesh_kind, a, s.wave, n, s.stats_level >= 2 ? 1 : 0, s.stream), "render_wave_kernel");
      launches += 2;
      s.spp += n * per;                                                  // ray_tracer.adb:168-176
    }
    samples = pixels * (uint64_t)(units * per);
  }
  CU(cudaEventRecord(s.ev1, s.stream), "event");
  if (stats) {
    unsigned long long c[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(c, s.d_rays, sizeof(c), cudaMemcpyDeviceToHost, s.stream), "memcpy counters");
    CU(cudaStreamSynchronize(s.stream), "render sync");
    float ms = 0.f; cudaEventElapsedTime(&ms, s.ev0, s.ev1);
    stats->rays = c[0] - s.last_counters[0]; stats->mesh_queries = c[1] - s.last_counters[1]; stats->bounces = c[2] - s.last_counters[2];
    stats->reserved0 = 0;
    s.last_counters[0] = c[0]; s.last_counters[1] = c[1]; s.last_counters[2] = c[2]; s.last_counters[3] = c[3];
    stats->samples = samples; stats->kernel_ms = ms; stats->total_ms = ms; stats->spp = s.spp;
    stats->kernel_launches = launches;
  }
  return 0;
}

// normC = 1 / g_spp for path-traced frames; the debug palette is resolved as it is (ray_tracer.adb:245-257 vs :281-291)
int resolve_spp() {
  State& s = S();
  return (s.palette_frame || s.spp <= 0) ? 1 : s.spp;
}

int read_frame_locked(float* hdr4, uint32_t* ldr, int* launches) {
  State& s = S();
  if (!s.d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (s.world != 1) return fail(GCORE_E_STATE, "gcore_read_frame needs world == 1; gather the accumulators and use gcore_resolve_gathered_device");
  if (!hdr4 && !ldr) return 0;
  if (int rc = ensure_out_buffers(hdr4 != nullptr, ldr != nullptr)) return rc;
  RenderArgs a = make_args();
  CU(gcore_dev::launch_resolve(a, resolve_spp(), hdr4 ? s.d_hdr : nullptr, ldr ? s.d_ldr : nullptr, s.stream), "resolve_kernel");
  if (launches) *launches += 1;
  size_t px = (size_t)s.width * s.height;
  if (hdr4) CU(cudaMemcpyAsync(hdr4, s.d_hdr, sizeof(float4) * px, cudaMemcpyDeviceToHost, s.stream), "memcpy hdr");
  if (ldr) CU(cudaMemcpyAsync(ldr, s.d_ldr, sizeof(uint32_t) * px, cudaMemcpyDeviceToHost, s.stream), "memcpy ldr");
  CU(cudaStreamSynchronize(s.stream), "read_frame sync");
  return 0;
}

}  // namespace

// ======================================================================================================================
extern "C" {

// ---------------------------------------------------------------------------------------------- diagnostics
int gcore_last_error(char* buf, int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (buf && n > 0) { strncpy(buf, s.err_text.c_str(), (size_t)n - 1); buf[n - 1] = 0; }
  return s.err_code;
}

int gcore_set_device(int device) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (s.dev_ready && device != s.device) return fail(GCORE_E_STATE, "gcore_set_device must be called before any other device work");
  s.device = device;
  return ensure_device();
}

int gcore_set_partition(int rank, int world) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (world < 1 || rank < 0 || rank >= world) return fail(GCORE_E_ARG, "bad partition");
  s.rank = rank; s.world = world;
  if (s.width > 0) return alloc_accum();
  return 0;
}

int gcore_stream(void** stream) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = ensure_device()) return rc;
  if (stream) *stream = (void*)s.stream;
  return 0;
}

int gcore_synchronize(void) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = ensure_device()) return rc;
  CU(cudaStreamSynchronize(s.stream), "synchronize");
  return 0;
}

int gcore_bvh_info(int64_t* ntris, int64_t* nnodes, int64_t* bytes, float* build_ms) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  const LbvhBuffers& b = s.mesh_bvh.ntris ? s.mesh_bvh : s.legacy_bvh;
  if (ntris) *ntris = b.ntris;
  if (nnodes) *nnodes = b.nnodes;
  if (bytes) *bytes = (int64_t)b.bytes;
  if (build_ms) *build_ms = b.build_ms;
  return 0;
}

int gcore_kernel_regs(int render_type) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = ensure_device()) return rc;
  return gcore_dev::render_kernel_regs(render_type);
}

// ---------------------------------------------------------------------------------------------- legacy seam
void gcore_destroy(void) {                                             // embree_connect.cpp:51-58
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  s.meshes.clear(); s.instances.clear();
  if (s.dev_ready) free_legacy_device();
  s.legacy_alive = false;
}

void gcore_init_and_clear(void) {                                      // embree_connect.cpp:60-67
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  gcore_destroy();
  s.legacy_alive = true;
}

int gcore_add_mesh_3f(const float* a_vertices3f, int a_vertexNum, const int* a_indices, int a_indicesNum) {   // :69-144
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!a_vertices3f || !a_indices || a_vertexNum <= 0 || a_indicesNum < 0) { fail(GCORE_E_ARG, "gcore_add_mesh_3f: bad arguments"); return -1; }
  LegacyMesh m;
  const int triNum = a_indicesNum / 3;
  m.idx.resize(3 * (size_t)triNum);
  int maxVertexId = 0;
  for (int i = 0; i < triNum; i++) {
    int A = a_indices[i * 3 + 0], B = a_indices[i * 3 + 1], C = a_indices[i * 3 + 2];
    if (A >= a_vertexNum || B >= a_vertexNum || C >= a_vertexNum || A < 0 || B < 0 || C < 0) A = B = C = a_vertexNum - 1;   // :98-109
    if (A > maxVertexId) maxVertexId = A;
    if (B > maxVertexId) maxVertexId = B;
    if (C > maxVertexId) maxVertexId = C;
    m.idx[3 * (size_t)i] = A; m.idx[3 * (size_t)i + 1] = B; m.idx[3 * (size_t)i + 2] = C;
  }
  // the shim copies `maxVertexId` vertices (:122-134), one short of what the indices reference; copy maxVertexId + 1
  m.verts.assign(a_vertices3f, a_vertices3f + 3 * (size_t)(maxVertexId + 1));
  s.meshes.push_back(std::move(m));
  return (int)s.meshes.size() - 1;
}

void gcore_instance_meshes(int a_geomId, const float* a_matrices16f, int a_matrixNum) {   // :147-184
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  s.err_code = 0; s.err_text.clear();
  if (a_geomId < 0 || a_geomId >= (int)s.meshes.size() || !a_matrices16f) { fail(GCORE_E_ARG, "gcore_instance_meshes: bad meshId"); return; }
  for (int k = 0; k < a_matrixNum; k++) {
    LegacyInstance in; in.mesh = a_geomId;
    memcpy(in.m, a_matrices16f + 16 * (size_t)k, 12 * sizeof(float));   // RTC_FORMAT_FLOAT3X4_ROW_MAJOR (:169)
    s.instances.push_back(in);
  }
}

void gcore_commit_scene(void) {                                        // :241-244
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  s.err_code = 0; s.err_text.clear();                                  // void signature: gcore_last_error() afterwards speaks for this call
  if (ensure_device()) return;
  free_legacy_device();
  std::vector<LegacyInstance> inst = s.instances;
  if (inst.empty()) {                                                  // no instance declared: every mesh once, identity
    for (int m = 0; m < (int)s.meshes.size(); m++) { LegacyInstance in; in.mesh = m; memset(in.m, 0, sizeof(in.m)); in.m[0] = in.m[5] = in.m[10] = 1.0f; inst.push_back(in); }
  }
  std::vector<float> verts; std::vector<int> tris;
  for (int ii = 0; ii < (int)inst.size(); ii++) {
    const LegacyMesh& m = s.meshes[inst[ii].mesh];
    const float* M = inst[ii].m;
    int base = (int)verts.size() / 3;
    for (size_t v = 0; v + 2 < m.verts.size(); v += 3) {
      float x = m.verts[v], y = m.verts[v + 1], z = m.verts[v + 2];
      verts.push_back(M[0] * x + M[1] * y + M[2] * z + M[3]);
      verts.push_back(M[4] * x + M[5] * y + M[6] * z + M[7]);
      verts.push_back(M[8] * x + M[9] * y + M[10] * z + M[11]);
    }
    for (size_t t = 0; t + 2 < m.idx.size(); t += 3) {
      tris.push_back(base + m.idx[t]); tris.push_back(base + m.idx[t + 1]); tris.push_back(base + m.idx[t + 2]);
      s.h_legacy_prim.push_back((int)(t / 3)); s.h_legacy_inst.push_back(ii);
    }
  }
  const int ntris = (int)tris.size() / 3;
  if (ntris == 0) return;
  cudaError_t e;
  if ((e = cudaMalloc(&s.d_legacy_verts, verts.size() * 4)) != cudaSuccess || (e = cudaMalloc(&s.d_legacy_tris, tris.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&s.d_legacy_prim, 4 * (size_t)ntris)) != cudaSuccess || (e = cudaMalloc(&s.d_legacy_inst, 4 * (size_t)ntris)) != cudaSuccess) {
    fail_cuda(e, "cudaMalloc legacy scene"); free_legacy_device(); return;
  }
  cudaMemcpyAsync(s.d_legacy_verts, verts.data(), verts.size() * 4, cudaMemcpyHostToDevice, s.stream);
  cudaMemcpyAsync(s.d_legacy_tris, tris.data(), tris.size() * 4, cudaMemcpyHostToDevice, s.stream);
  cudaMemcpyAsync(s.d_legacy_prim, s.h_legacy_prim.data(), 4 * (size_t)ntris, cudaMemcpyHostToDevice, s.stream);
  cudaMemcpyAsync(s.d_legacy_inst, s.h_legacy_inst.data(), 4 * (size_t)ntris, cudaMemcpyHostToDevice, s.stream);
  e = gcore_dev::lbvh_build(s.d_legacy_verts, (int)verts.size() / 3, s.d_legacy_tris, ntris, nullptr, &s.legacy_bvh, s.stream);
  if (e != cudaSuccess) { fail_cuda(e, "lbvh_build (legacy scene)"); free_legacy_device(); return; }
  s.scene_dirty = true;
}

// device staging of the batched seam queries: grown on demand and kept, so that the one-ray-per-call legacy entry point
// (scene_hydra_embree.adb:426-446 calls it for every ray) costs a launch and two small copies, not six allocations
static int ensure_batch_buffers(size_t n) {
  State& s = S();
  if (n <= s.batch_cap) return 0;
  cudaFree(s.d_batch_rays); cudaFree(s.d_batch_hits); cudaFree(s.d_batch_found);
  s.d_batch_rays = nullptr; s.d_batch_hits = nullptr; s.d_batch_found = nullptr; s.batch_cap = 0;
  size_t cap = n < 256 ? 256 : n;
  CU(cudaMalloc(&s.d_batch_rays, sizeof(gcore_ray) * cap), "cudaMalloc batch");
  CU(cudaMalloc(&s.d_batch_hits, sizeof(gcore_HitCpp) * cap), "cudaMalloc batch");
  CU(cudaMalloc(&s.d_batch_found, sizeof(int) * cap), "cudaMalloc batch");
  s.batch_cap = cap;
  return 0;
}

int gcore_bvh_closest_hit_batch(int n, const gcore_ray* rays, float t_near, float t_far, gcore_HitCpp* hits, int32_t* found) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (n < 0 || (n > 0 && (!rays || !hits || !found))) return fail(GCORE_E_ARG, "gcore_bvh_closest_hit_batch: bad arguments");
  if (n == 0) return 0;
  if (int rc = sync_scene()) return rc;
  if (int rc = ensure_batch_buffers((size_t)n)) return rc;
  gcore_ray* d_r = s.d_batch_rays; gcore_HitCpp* d_h = s.d_batch_hits; int* d_f = s.d_batch_found;
  cudaError_t e = cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_h, hits, sizeof(gcore_HitCpp) * (size_t)n, cudaMemcpyHostToDevice, s.stream);   // records are only written on a hit
  if (e == cudaSuccess) e = gcore_dev::launch_bvh_batch(n, d_r, t_near, t_far, d_h, d_f, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_HitCpp) * (size_t)n, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(found, d_f, 4 * (size_t)n, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) return fail_cuda(e, "bvh batch");
  // flattened triangle id -> (primIndex inside its mesh, instIndex); Ng from the flattened world-space triangle
  for (int i = 0; i < n; i++) if (found[i]) {
    int flat = hits[i].primIndex;
    if (flat >= 0 && flat < (int)s.h_legacy_prim.size()) { hits[i].primIndex = s.h_legacy_prim[flat]; hits[i].instIndex = s.h_legacy_inst[flat]; hits[i].geomIndex = 0; }
  }
  return 0;
}

int gcore_bvh_bench(int n, const gcore_ray* rays, float t_near, float t_far, int which, int repeat, float* ms_per_launch, uint64_t* stats3) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (n <= 0 || !rays || repeat < 1 || which < 0 || which > 1) return fail(GCORE_E_ARG, "gcore_bvh_bench: bad arguments");
  if (int rc = sync_scene()) return rc;
  if ((which ? s.mesh_bvh.ntris : s.legacy_bvh.ntris) <= 0) return fail(GCORE_E_STATE, "gcore_bvh_bench: no LBVH of that kind is built");
  gcore_ray* d_r = nullptr; gcore_HitCpp* d_h = nullptr; int* d_f = nullptr; unsigned long long* d_c = nullptr;
  auto cleanup = [&]() { cudaFree(d_r); cudaFree(d_h); cudaFree(d_f); cudaFree(d_c); };
  cudaError_t e;
  if ((e = cudaMalloc(&d_r, sizeof(gcore_ray) * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_h, sizeof(gcore_HitCpp) * (size_t)n)) != cudaSuccess ||
      (e = cudaMalloc(&d_f, 4 * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_c, 24)) != cudaSuccess) { cleanup(); return fail_cuda(e, "cudaMalloc bvh bench"); }
  cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, s.stream);
  cudaMemsetAsync(d_c, 0, 24, s.stream);
  e = gcore_dev::launch_bvh_batch_which(n, d_r, t_near, t_far, d_h, d_f, which, d_c, s.stream);          // counting pass (also the warm-up)
  unsigned long long c[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(c, d_c, 24, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaEventRecord(s.ev0, s.stream);
  for (int k = 0; k < repeat && e == cudaSuccess; k++) e = gcore_dev::launch_bvh_batch_which(n, d_r, t_near, t_far, d_h, d_f, which, nullptr, s.stream);
  if (e == cudaSuccess) e = cudaEventRecord(s.ev1, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cleanup();
  if (e != cudaSuccess) return fail_cuda(e, "bvh bench");
  float ms = 0.f; cudaEventElapsedTime(&ms, s.ev0, s.ev1);
  if (ms_per_launch) *ms_per_launch = ms / (float)repeat;
  if (stats3) { stats3[0] = c[0]; stats3[1] = c[1]; stats3[2] = c[2]; }
  return 0;
}

bool gcore_closest_hit(const float a_rayPos[3], const float a_rayDir[3], float t_near, float t_far, gcore_HitCpp* pHit) {   // :196-238
  if (!a_rayPos || !a_rayDir || !pHit) return false;
  gcore_ray r; memcpy(r.origin, a_rayPos, 12); memcpy(r.direction, a_rayDir, 12);
  gcore_HitCpp h = *pHit; int32_t f = 0;
  if (gcore_bvh_closest_hit_batch(1, &r, t_near, t_far, &h, &f) != 0) return false;
  if (f) *pHit = h;                                                    // written only on a hit (:222-235)
  return f != 0;
}

// ---------------------------------------------------------------------------------------------- analytic scene
int gcore_set_analytic_scene(const gcore_sphere* spheres, int nspheres, const gcore_cornell* box, const gcore_flatlight* fl) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (nspheres < 0 || nspheres > GC_MAX_SPHERES || (nspheres > 0 && !spheres)) return fail(GCORE_E_LIMIT, "gcore_set_analytic_scene: 0..96 spheres");
  s.hs.nspheres = nspheres;
  for (int i = 0; i < nspheres; i++) {
    s.hs.sph[i][0] = spheres[i].pos[0]; s.hs.sph[i][1] = spheres[i].pos[1]; s.hs.sph[i][2] = spheres[i].pos[2]; s.hs.sph[i][3] = spheres[i].r;
    s.hs.sph_mat[i] = spheres[i].mat;
  }
  if (box) {
    memcpy(s.hs.box_min, box->bmin, 12); memcpy(s.hs.box_max, box->bmax, 12);
    memcpy(s.hs.box_mat, box->mat_indices, 24); memcpy(s.hs.box_nrm, box->normals, 72);
  }
  if (fl) { memcpy(s.hs.fl_min, fl->boxMin, 12); memcpy(s.hs.fl_max, fl->boxMax, 12); }
  s.scene_dirty = true;
  return 0;
}

int gcore_set_materials(const gcore_material* materials, int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (n < 0 || n > GC_MAX_MATERIALS - 1 || (n > 0 && !materials)) return fail(GCORE_E_LIMIT, "gcore_set_materials: 0..31 materials");
  memset(s.hs.mats, 0, sizeof(s.hs.mats));
  for (int i = 0; i < n; i++) {
    art::DMaterial& d = s.hs.mats[i];
    d.type = materials[i].type; memcpy(d.color, materials[i].color, 12); memcpy(d.color2, materials[i].color2, 12);
    d.param = materials[i].param; d.light = materials[i].light;
  }
  s.hs.nmaterials = n;
  s.scene_dirty = true;
  return 0;
}

int gcore_set_light(const gcore_light* l) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!l) return fail(GCORE_E_ARG, "gcore_set_light: null");
  s.hs.light_shape = l->shape;
  memcpy(s.hs.l_boxMin, l->boxMin, 12); memcpy(s.hs.l_boxMax, l->boxMax, 12); memcpy(s.hs.l_normal, l->normal, 12);
  memcpy(s.hs.l_center, l->center, 12); s.hs.l_radius = l->radius; memcpy(s.hs.l_intensity, l->intensity, 12);
  s.hs.l_area = l->surfaceArea;
  s.scene_dirty = true;
  return 0;
}

int gcore_set_camera(const gcore_camera* c) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!c) return fail(GCORE_E_ARG, "gcore_set_camera: null");
  memcpy(s.hs.cam_pos, c->pos, 12);
  memcpy(s.hs.cam_m, c->matrix, 48);                                   // rows 0..2 of the row-major 4x4
  s.scene_dirty = true;
  return 0;
}

int gcore_set_mesh(const float* pos3, const float* nrm3, int nverts, const int32_t* tris3, int ntris, const float* bbox6, int mesh_mode) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (nverts < 0 || ntris < 0 || (ntris > 0 && (!pos3 || !nrm3 || !tris3))) return fail(GCORE_E_ARG, "gcore_set_mesh: bad arguments");
  if (mesh_mode != GCORE_MESH_REFERENCE_BF && mesh_mode != GCORE_MESH_NEAREST) return fail(GCORE_E_ARG, "gcore_set_mesh: bad mesh_mode");
  s.mesh_pos.assign(pos3, pos3 + 3 * (size_t)nverts);
  s.mesh_nrm.assign(nrm3, nrm3 + 3 * (size_t)nverts);
  s.mesh_tris.assign(tris3, tris3 + 3 * (size_t)ntris);
  if (bbox6) memcpy(s.mesh_bbox, bbox6, 24);
  else { memset(s.mesh_bbox, 0, 24); fold_bbox(s.mesh_pos, s.mesh_bbox); }
  return install_mesh(mesh_mode);
}

int gcore_set_mesh_vsgf(const char* path, const float* matrix16, int mesh_mode) {      // geometry.adb:499-609
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!path || !matrix16) return fail(GCORE_E_ARG, "gcore_set_mesh_vsgf: null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(GCORE_E_IO, std::string("cannot open ") + path);
  struct { int64_t fileSizeInBytes; int32_t verticesNum, indicesNum, materialsNum, flags; } hd;   // :499-507
  if (fread(&hd, 24, 1, f) != 1 || hd.verticesNum <= 0 || hd.indicesNum < 0) { fclose(f); return fail(GCORE_E_IO, "bad .vsgf header"); }
  const int vn = hd.verticesNum, tn = hd.indicesNum / 3;
  std::vector<float> p4(4 * (size_t)vn), n4(4 * (size_t)vn), skip(4 * (size_t)vn);
  std::vector<int> idx(3 * (size_t)tn), mids(tn);
  bool ok = fread(p4.data(), 16, vn, f) == (size_t)vn && fread(n4.data(), 16, vn, f) == (size_t)vn &&
            fread(skip.data(), 8, vn, f) == (size_t)vn;                                 // uvs are read and zeroed (:562-567)
  if (ok && hd.flags != 0) ok = fread(skip.data(), 16, vn, f) == (size_t)vn;              // tangents skipped (:571-575)
  ok = ok && fread(idx.data(), 12, tn, f) == (size_t)tn && fread(mids.data(), 4, tn, f) == (size_t)tn;
  fclose(f);
  if (!ok) return fail(GCORE_E_IO, "truncated .vsgf file");
  s.mesh_pos.resize(3 * (size_t)vn); s.mesh_nrm.resize(3 * (size_t)vn);
  const float* M = matrix16;
  for (int i = 0; i < vn; i++) {
    float x = p4[4 * (size_t)i], y = p4[4 * (size_t)i + 1], z = p4[4 * (size_t)i + 2];
    s.mesh_pos[3 * (size_t)i]     = M[0] * x + M[1] * y + M[2] * z + M[3];               // vector_math.adb:137-144
    s.mesh_pos[3 * (size_t)i + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    s.mesh_pos[3 * (size_t)i + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
    s.mesh_nrm[3 * (size_t)i] = n4[4 * (size_t)i]; s.mesh_nrm[3 * (size_t)i + 1] = n4[4 * (size_t)i + 1]; s.mesh_nrm[3 * (size_t)i + 2] = n4[4 * (size_t)i + 2];
  }
  s.mesh_tris = idx;
  memset(s.mesh_bbox, 0, 24);
  fold_bbox(s.mesh_pos, s.mesh_bbox);
  return install_mesh(mesh_mode);
}

int gcore_scene_init_cornell(const char* vsgf_path) {                                     // scene.adb:89-217
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  const float M_PI_F = 3.14159265358979323846f;
  gcore_light l; memset(&l, 0, sizeof(l));
  l.shape = GCORE_LIGHT_SPHERE;                                                          // g_lightRef := sphLight (:128)
  l.boxMin[0] = -0.75f; l.boxMin[1] = 4.98f; l.boxMin[2] = 1.25f; l.boxMax[0] = 0.75f; l.boxMax[1] = 4.98f; l.boxMax[2] = 3.25f;
  l.normal[1] = -1.0f;
  l.center[0] = 0.0f; l.center[1] = 4.5f; l.center[2] = 1.0f; l.radius = 0.5f;
  for (int k = 0; k < 3; k++) l.intensity[k] = 0.5f * 20.0f;                             // intensity*0.5 (:121)
  l.surfaceArea = 4.0f * M_PI_F * l.radius * l.radius;                                   // :122
  gcore_sphere sph[3] = {{{-1.5f, 1.0f, 1.5f}, 1.0f, 8}, {{1.4f, 1.0f, 3.0f}, 1.0f, 0}, {{0.0f, 4.5f, 1.0f}, 0.5f, 4}};   // :139-144,182-192
  gcore_material mats[11];
  for (auto& m : mats) { memset(&m, 0, sizeof(m)); m.light = -1; }
  auto setm = [&](int i, int type, float r, float g, float b, float r2, float g2, float b2, float p, int light) {
    mats[i].type = type; mats[i].color[0] = r; mats[i].color[1] = g; mats[i].color[2] = b;
    mats[i].color2[0] = r2; mats[i].color2[1] = g2; mats[i].color2[2] = b2; mats[i].param = p; mats[i].light = light;
  };
  setm(0, GCORE_MAT_GLASS, 0.75f, 0.75f, 0.75f, 0.85f, 0.85f, 0.85f, 1.75f, -1);        // :153-185
  setm(1, GCORE_MAT_LAMBERT, 0.5f, 0.5f, 0.5f, 0, 0, 0, 0, -1);
  setm(2, GCORE_MAT_LAMBERT, 0.25f, 0.5f, 0.0f, 0, 0, 0, 0, -1);
  setm(3, GCORE_MAT_LAMBERT, 0.5f, 0.0f, 0.0f, 0, 0, 0, 0, -1);
  setm(4, GCORE_MAT_LIGHT, 0, 0, 0, 0, 0, 0, 0, 0);
  setm(5, GCORE_MAT_MIRROR, 0.75f, 0.75f, 0.75f, 0, 0, 0, 0, -1);
  setm(8, GCORE_MAT_PHONG, 0.75f, 0.75f, 0.75f, 0, 0, 0, 80.0f, -1);
  mats[9] = mats[1]; mats[10] = mats[1];
  gcore_cornell box = {{-2.5f, 0.0f, 0.0f}, {2.5f, 5.0f, 5.0f}, {2, 3, 1, 1, 8, 1},       // scene.ads:75-80
                       {1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1}};
  gcore_camera cam; memset(&cam, 0, sizeof(cam));                                        // :212-215
  cam.pos[0] = 0.0f; cam.pos[1] = 2.55f; cam.pos[2] = 12.5f; cam.up[1] = 1.0f;
  cam.matrix[0] = cam.matrix[5] = cam.matrix[10] = cam.matrix[15] = 1.0f;
  int rc;
  if ((rc = gcore_set_light(&l))) return rc;
  if ((rc = gcore_set_analytic_scene(sph, 3, &box, nullptr))) return rc;
  if ((rc = gcore_set_materials(mats, 11))) return rc;
  if ((rc = gcore_set_camera(&cam))) return rc;
  M4 mrot = rotation(-M_PI_F / 6.0f, 0.0f, 1.0f, 0.0f);                                  // :194-206
  M4 mscale = ident(), mtans = ident();
  mtans.m[0][3] = -0.75f; mtans.m[1][3] = 0.1f; mtans.m[2][3] = 3.1f; mtans.m[3][3] = 1.0f;
  mscale.m[0][0] = 2.0f; mscale.m[1][1] = 2.0f; mscale.m[2][2] = 2.0f;
  M4 M = mul(mul(mtans, mrot), mscale);
  return gcore_set_mesh_vsgf(vsgf_path, &M.m[0][0], GCORE_MESH_REFERENCE_BF);
}

int gcore_get_mesh(float* pos3, float* nrm3, int32_t* tris3, float* bbox6, int* nverts, int* ntris) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (pos3 && !s.mesh_pos.empty()) memcpy(pos3, s.mesh_pos.data(), s.mesh_pos.size() * 4);
  if (nrm3 && !s.mesh_nrm.empty()) memcpy(nrm3, s.mesh_nrm.data(), s.mesh_nrm.size() * 4);
  if (tris3 && !s.mesh_tris.empty()) memcpy(tris3, s.mesh_tris.data(), s.mesh_tris.size() * 4);
  if (bbox6) memcpy(bbox6, s.mesh_bbox, 24);
  if (nverts) *nverts = (int)s.mesh_pos.size() / 3;
  if (ntris) *ntris = (int)s.mesh_tris.size() / 3;
  return 0;
}

// ---------------------------------------------------------------------------------------------- batched scene queries
int gcore_closest_hit_batch(int n, const gcore_ray* rays, gcore_hit* hits) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (n < 0 || (n > 0 && (!rays || !hits))) return fail(GCORE_E_ARG, "gcore_closest_hit_batch: bad arguments");
  if (n == 0) return 0;
  if (int rc = sync_scene()) return rc;
  gcore_ray* d_r = nullptr; gcore_hit* d_h = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d_r, sizeof(gcore_ray) * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_h, sizeof(gcore_hit) * (size_t)n)) != cudaSuccess) {
    cudaFree(d_r); cudaFree(d_h); return fail_cuda(e, "cudaMalloc batch");
  }
  cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, s.stream);
  e = gcore_dev::launch_closest_batch(s.hs.mesh_kind, n, d_r, d_h, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_hit) * (size_t)n, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cudaFree(d_r); cudaFree(d_h);
  if (e != cudaSuccess) return fail_cuda(e, "closest batch");
  return 0;
}

// ---------------------------------------------------------------------------------------------- rendering
int gcore_init_render(const gcore_render_config* cfg) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!cfg) return fail(GCORE_E_ARG, "gcore_init_render: null");
  if (cfg->render_type < GCORE_RT_DEBUG || cfg->render_type > GCORE_PT_MIS) return fail(GCORE_E_ARG, "bad render_type");
  if (cfg->max_depth < 1 || cfg->max_depth > 16) return fail(GCORE_E_ARG, "max_depth must be in 1..16");
  s.cfg = *cfg;
  return 0;
}

int gcore_resize_viewport(int width, int height) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (width <= 0 || height <= 0 || (int64_t)width * height > (int64_t)1 << 28) return fail(GCORE_E_ARG, "bad viewport");
  s.width = width; s.height = height;
  return alloc_accum();
}

int gcore_render_units(int units, gcore_render_stats* stats) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  return render_units_locked(units, stats);
}

int gcore_get_counters(uint64_t* out3) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!out3) return fail(GCORE_E_ARG, "gcore_get_counters: null");
  if (int rc = ensure_device()) return rc;
  unsigned long long c[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(c, s.d_rays, sizeof(c), cudaMemcpyDeviceToHost, s.stream), "memcpy counters");
  CU(cudaStreamSynchronize(s.stream), "counters sync");
  out3[0] = c[0]; out3[1] = c[1]; out3[2] = c[2]; out3[3] = c[3];
  return 0;
}

int gcore_set_stats_level(int level) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (level < 1 || level > 2) return fail(GCORE_E_ARG, "gcore_set_stats_level: 1 or 2");
  s.stats_level = level;
  return 0;
}

int gcore_get_bvh_counters(uint64_t* out3) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!out3) return fail(GCORE_E_ARG, "gcore_get_bvh_counters: null");
  if (int rc = ensure_device()) return rc;
  unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
  CU(cudaMemcpyAsync(c, s.d_rays, sizeof(c), cudaMemcpyDeviceToHost, s.stream), "memcpy counters");
  CU(cudaStreamSynchronize(s.stream), "counters sync");
  out3[0] = c[3]; out3[1] = c[4]; out3[2] = c[5];
  return 0;
}

int gcore_get_spp(void) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  return s.spp;
}

int gcore_read_frame(float* hdr4, uint32_t* ldr) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  return read_frame_locked(hdr4, ldr, nullptr);
}

int gcore_render(const gcore_render_config* cfg, int width, int height, int units, float* hdr4, uint32_t* ldr, gcore_render_stats* stats) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  int rc;
  if ((rc = gcore_init_render(cfg))) return rc;
  if (width != s.width || height != s.height || !s.d_accum) { if ((rc = gcore_resize_viewport(width, height))) return rc; }
  else {                                                               // same viewport: clear like Resize_Viewport does
    CU(cudaMemsetAsync(s.d_accum, 0, sizeof(float4) * (size_t)s.nslots, s.stream), "memset accum");
    CU(cudaMemsetAsync(s.d_rays, 0, 6 * sizeof(unsigned long long), s.stream), "memset counters");
    s.last_counters[0] = s.last_counters[1] = s.last_counters[2] = s.last_counters[3] = 0;
    s.spp = 0; s.palette_frame = false;
  }
  gcore_render_stats st; memset(&st, 0, sizeof(st));
  if ((rc = render_units_locked(units, &st))) return rc;
  int launches = st.kernel_launches;
  if ((rc = read_frame_locked(hdr4, ldr, &launches))) return rc;
  CU(cudaEventRecord(s.ev2, s.stream), "event");
  CU(cudaEventSynchronize(s.ev2), "event sync");
  float ms = 0.f; cudaEventElapsedTime(&ms, s.ev0, s.ev2);
  st.total_ms = ms; st.kernel_launches = launches;
  if (stats) *stats = st;
  return 0;
}

int gcore_load_frame(const float* hdr4, int spp) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!hdr4 || spp < 0) return fail(GCORE_E_ARG, "gcore_load_frame: bad arguments");
  if (!s.d_accum) return fail(GCORE_E_STATE, "gcore_load_frame: no viewport");
  if (int rc = ensure_out_buffers(true, false)) return rc;
  size_t px = (size_t)s.width * s.height;
  CU(cudaMemcpyAsync(s.d_hdr, hdr4, sizeof(float4) * px, cudaMemcpyHostToDevice, s.stream), "memcpy hdr");
  RenderArgs a = make_args();
  CU(gcore_dev::launch_load_accum(a, s.d_hdr, s.stream), "load_accum_kernel");
  CU(cudaStreamSynchronize(s.stream), "load_frame sync");
  s.spp = spp; s.palette_frame = false;
  return 0;
}

int gcore_primary_hits(int width, int height, gcore_hit* hits) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (width <= 0 || height <= 0 || !hits) return fail(GCORE_E_ARG, "gcore_primary_hits: bad arguments");
  if (int rc = sync_scene()) return rc;
  const int sw = s.width, sh = s.height, srank = s.rank, sworld = s.world, sslots = s.nslots;
  s.width = width; s.height = height; s.rank = 0; s.world = 1;
  RenderArgs a = make_args();
  s.width = sw; s.height = sh; s.rank = srank; s.world = sworld;
  a.nslots = a.ntiles * GC_TILE * GC_TILE; a.accum = nullptr;
  (void)sslots;
  gcore_hit* d_h = nullptr;
  size_t px = (size_t)width * height;
  CU(cudaMalloc(&d_h, sizeof(gcore_hit) * px), "cudaMalloc hits");
  cudaError_t e = gcore_dev::launch_primary_hits(s.hs.mesh_kind, a, d_h, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_hit) * px, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cudaFree(d_h);
  if (e != cudaSuccess) return fail_cuda(e, "primary hits");
  return 0;
}

int gcore_accum_device_ptr(void** d_accum, int64_t* nslots) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!s.d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (d_accum) *d_accum = s.d_accum;
  if (nslots) *nslots = s.nslots;
  return 0;
}

int gcore_resolve_device(void* d_hdr4, void* d_ldr) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!s.d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (s.world != 1) return fail(GCORE_E_STATE, "gcore_resolve_device needs world == 1");
  RenderArgs a = make_args();
  CU(gcore_dev::launch_resolve(a, resolve_spp(), (float4*)d_hdr4, (uint32_t*)d_ldr, s.stream), "resolve_kernel");
  return 0;
}

int gcore_resolve_gathered_device(const void* d_all_accum, int spp, void* d_hdr4, void* d_ldr) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!d_all_accum || s.width <= 0) return fail(GCORE_E_ARG, "gcore_resolve_gathered_device: bad arguments");
  if (int rc = ensure_device()) return rc;
  RenderArgs a = make_args();
  CU(gcore_dev::launch_resolve_gathered(a, (const float4*)d_all_accum, spp > 0 ? spp : 1, (float4*)d_hdr4, (uint32_t*)d_ldr, s.stream), "resolve_kernel");
  return 0;
}

}  // extern "C"
