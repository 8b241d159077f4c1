// gcore_api.cu -- the C ABI of libgcore_b200.so (include/gcore.h): process-wide render-core state, scene marshalling,
// kernel launches.  Replaces cpp/embree_connect.cpp (legacy seam) and the state of ray_tracer.adb / scene.adb that the
// hot path needs on the device.  No CPU compute path exists here: every query / render call launches CUDA kernels.
//
// One host process drives 1..8 GPUs (the reference host is ONE Ada process, ray_tracer.adb:264-277): the scene description
// lives once on the host, every device holds its own copy (constant block + LBVHs, rebuilt deterministically per device),
// renders the tiles t with t % G == d into its own slot-major accumulator, and the accumulators meet on device 0 through
// one grouped ncclSend/ncclRecv per frame read.
#include "gcore_internal.h"
#include "art_types.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

using gcore_dev::LbvhBuffers;
using gcore_dev::RenderArgs;
using gcore_dev::WaveBuffers;

namespace {

#define GC_MAX_GPUS 8

struct LegacyMesh { std::vector<float> verts; std::vector<int> idx; std::vector<float> normals; int mat = 2; };
struct LegacyInstance { int mesh; float m[12]; };

// the committed legacy scene, flattened to world space on the host (every device builds its LBVH from this)
struct FlatScene {
  std::vector<float> verts; std::vector<int> tris;
  std::vector<int> prim, inst, mat;          // flattened triangle -> triangle index inside its mesh / instance id / material id
  std::vector<float> n9;                     // flattened triangle -> nA, nB, nC (vertex normals under the instance's 3x3, or the geometric normal)
  float tight_min[3] = {0, 0, 0}, tight_max[3] = {0, 0, 0};
  int version = 0;
};

// NCCL entry points, bound at run time: a host process that already carries an NCCL (PyTorch bundles its own libnccl.so.2)
// must not get a second one through our link line
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Device {
  int  ordinal = 0;
  bool ready = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  bool scene_dirty = true;
  int  mesh_version = -1, legacy_version = -1;      // what mesh_bvh / legacy_bvh were built from
  LbvhBuffers mesh_bvh, legacy_bvh;
  float4* d_accum = nullptr;                        // device 0: the first nslots of d_gather
  float4* d_gather = nullptr;                       // device 0 only: gpu_count accumulators back to back
  int nslots = 0;
  unsigned long long* d_rays = nullptr;             // 6 cumulative counters (RenderArgs::counters)
  unsigned long long last_counters[6] = {0, 0, 0, 0, 0, 0};
  float last_kernel_ms = 0.f;
  float4* d_hdr = nullptr; uint32_t* d_ldr = nullptr; size_t out_pixels = 0;
  gcore_ray* d_batch_rays = nullptr; gcore_HitCpp* d_batch_hits = nullptr; int* d_batch_found = nullptr; size_t batch_cap = 0;
  WaveBuffers wave;
  int sm_count = 0;
};

struct State {
  std::recursive_mutex mu;
  int  base_device = 0, gpu_count = 1;
  Device dev[GC_MAX_GPUS];
  NcclApi nccl; ncclComm_t comms[GC_MAX_GPUS]; bool nccl_ready = false;

  // legacy seam
  bool legacy_alive = false;
  std::vector<LegacyMesh> meshes;
  std::vector<LegacyInstance> instances;
  FlatScene flat;
  bool use_committed = false;                // render source: false = gcore_set_mesh mesh, true = the committed legacy scene
  int  seam_two_sided = 0;

  // analytic scene (host mirror of the __constant__ block)
  art::DeviceScene hs;
  std::vector<float> mesh_pos, mesh_nrm; std::vector<int> mesh_tris; float mesh_bbox[6] = {0, 0, 0, 0, 0, 0};
  int mesh_version = 0;

  // render state
  gcore_render_config cfg{GCORE_PT_MIS, 1, 8, 0x5EED0003u, {0.f, 0.f, 0.f}};
  int width = 0, height = 0, spp = 0;
  bool palette_frame = false;                // the last pass was Debug_Ray_Tracing: resolve without 1/spp (ray_tracer.adb:245-257)
  int rank = 0, world = 1;                   // multi-PROCESS partition (one device per process); exclusive with gpu_count > 1
  int stats_level = 1;

  int  err_code = 0;
  std::string err_text;
};

State& S() { static State s; return s; }

// ---- the one-ray seam under many host threads (scene_hydra_embree.adb:426-446 calls gcore_closest_hit from every Ada task):
// flat combining.  A caller files its ray in the open batch; the first caller to find no leader becomes the leader, launches ONE
// kernel for everything filed so far, hands the answers out and goes on with the next batch while there is one.  T threads thus
// share a launch + synchronisation per ~T rays instead of queueing for one each.  Waiters spin on their own flag (a render task
// has nothing else to do, and a futex wake-up costs more than the launch); a new leader that knows of companions on their way in
// gives them up to 3 us to file before it launches.
// The seam's bookkeeping lock is held for a few dozen instructions at a time by up to Threads_Num callers per batch: a futex-backed mutex
// puts them to sleep and wakes them at 5-10 us apiece (measured: 49 instead of 22 us per batch at 4 callers), so it spins.
struct SpinLock {
  std::atomic<int> f{0};
  void lock() {
    int spins = 0;
    for (;;) {
      if (!f.exchange(1, std::memory_order_acquire)) return;
      while (f.load(std::memory_order_relaxed)) { if (++spins > 4096) { std::this_thread::yield(); spins = 0; } }
    }
  }
  void unlock() { f.store(0, std::memory_order_release); }
};

struct Seam {
  SpinLock m;
  gcore_dev::SeamBatch* h_req[2] = {nullptr, nullptr};         // mapped pinned host memory, two batches (one being filled, one in flight)
  gcore_dev::SeamBatch* d_req[2] = {nullptr, nullptr};         // the same memory as the device sees it
  struct Waiter { gcore_HitCpp* out; int found; std::atomic<int> done; };
  Waiter* waiter[2][GC_SEAM_BATCH];
  int  n_open = 0;                                             // requests filed in the open batch
  int  open_buf = 0;
  int  last_n = 1;                                             // size of the previous batch: how many companions to expect
  bool leader_active = false;
  std::atomic<int> inside{0};                                  // threads currently inside gcore_closest_hit
  unsigned long long launches = 0, rays = 0;                   // combined batches launched / rays answered (gcore_seam_stats)
  // the resident server (render_kernels.cu: bvh_seam_server); all of it is touched under State::mu only
  gcore_dev::SeamMailbox* h_mb = nullptr; gcore_dev::SeamMailbox* d_mb = nullptr;
  cudaStream_t srv_stream = nullptr;
  unsigned int seq = 0;                                        // 24-bit sequence number of the last batch posted
  int  srv_two_sided = 0;                                      // which instantiation is (or was last) resident
  bool srv_usable = true;                                      // false after a batch the server did not answer: back to one launch per batch
  unsigned long long srv_starts = 0, srv_batches = 0;
};
Seam& seam() { static Seam s; return s; }

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

// nothing throws across the ABI (include/gcore.h): every extern "C" body that can allocate runs inside one of these
#define GC_TRY try {
#define GC_CATCH_INT } catch (const std::bad_alloc&) { return fail(GCORE_E_LIMIT, "out of host memory"); } \
                       catch (const std::exception& e_) { return fail(GCORE_E_STATE, std::string("internal error: ") + e_.what()); } \
                       catch (...) { return fail(GCORE_E_STATE, "internal error"); }
#define GC_CATCH_VOID } catch (const std::bad_alloc&) { fail(GCORE_E_LIMIT, "out of host memory"); } \
                        catch (const std::exception& e_) { fail(GCORE_E_STATE, std::string("internal error: ") + e_.what()); } \
                        catch (...) { fail(GCORE_E_STATE, "internal error"); }

inline int world_eff() { State& s = S(); return s.world * s.gpu_count; }
inline int rank_eff(int d) { State& s = S(); return s.rank * s.gpu_count + d; }

int ensure_device(int d) {
  State& s = S();
  Device& dv = s.dev[d];
  if (dv.ready) { CU(cudaSetDevice(dv.ordinal), "cudaSetDevice"); return 0; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(GCORE_E_CUDA, "no CUDA device available (libgcore_b200 has no CPU fallback)"); }
  dv.ordinal = s.base_device + d;
  if (dv.ordinal >= n) return fail(GCORE_E_ARG, "device index out of range");
  CU(cudaSetDevice(dv.ordinal), "cudaSetDevice");
  CU(cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking), "cudaStreamCreate");
  CU(cudaEventCreate(&dv.ev0), "cudaEventCreate"); CU(cudaEventCreate(&dv.ev1), "cudaEventCreate"); CU(cudaEventCreate(&dv.ev2), "cudaEventCreate");
  CU(cudaMalloc(&dv.d_rays, 6 * sizeof(unsigned long long)), "cudaMalloc");
  CU(cudaMemset(dv.d_rays, 0, 6 * sizeof(unsigned long long)), "cudaMemset");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dv.ordinal), "cudaGetDeviceProperties");
  dv.sm_count = prop.multiProcessorCount;
  dv.ready = true;
  dv.scene_dirty = true;
  return 0;
}

void init_host_scene(art::DeviceScene& hs) {
  memset(&hs, 0, sizeof(hs));
  hs.light_shape = art::LIGHT_SPHERE;
  hs.cam_m[0] = hs.cam_m[5] = hs.cam_m[10] = 1.0f;
}

struct Init { Init() { init_host_scene(S().hs); } } g_init;

void mark_scene_dirty() { State& s = S(); for (int d = 0; d < GC_MAX_GPUS; d++) s.dev[d].scene_dirty = true; }

// ---- per-device copies of the two triangle scenes: built from the host data, lazily where needed, so that a device that
// joins later (gcore_set_gpu_count) or a scene that changes picks the right version up at its next use
int build_mesh_bvh(int d) {
  State& s = S(); Device& dv = s.dev[d];
  gcore_dev::lbvh_free(&dv.mesh_bvh);
  dv.mesh_version = s.mesh_version;
  const int ntris = (int)s.mesh_tris.size() / 3, nverts = (int)s.mesh_pos.size() / 3;
  if (ntris <= GC_MAX_BF_TRIS) return 0;                         // lives in the constant block
  auto N = [&](int v, int k) { return s.mesh_nrm[3 * (size_t)v + k]; };
  std::vector<float> n9(9 * (size_t)ntris);
  for (int t = 0; t < ntris; t++)
    for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) n9[9 * (size_t)t + 3 * c + k] = N(s.mesh_tris[3 * (size_t)t + c], k);
  float *d_v = nullptr, *d_n = nullptr; int* d_t = nullptr;
  auto cleanup = [&]() { cudaFree(d_v); cudaFree(d_n); cudaFree(d_t); };
  cudaError_t e;
  if ((e = cudaMalloc(&d_v, s.mesh_pos.size() * 4)) != cudaSuccess || (e = cudaMalloc(&d_t, s.mesh_tris.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&d_n, n9.size() * 4)) != cudaSuccess) { cleanup(); return fail_cuda(e, "cudaMalloc mesh"); }
  cudaMemcpyAsync(d_v, s.mesh_pos.data(), s.mesh_pos.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  cudaMemcpyAsync(d_t, s.mesh_tris.data(), s.mesh_tris.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  cudaMemcpyAsync(d_n, n9.data(), n9.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  e = gcore_dev::lbvh_build(d_v, nverts, d_t, ntris, d_n, nullptr, &dv.mesh_bvh, dv.stream);
  cudaStreamSynchronize(dv.stream);
  cleanup();
  if (e != cudaSuccess) { gcore_dev::lbvh_free(&dv.mesh_bvh); return fail_cuda(e, "lbvh_build"); }
  return 0;
}

int build_legacy_bvh(int d) {
  State& s = S(); Device& dv = s.dev[d];
  gcore_dev::lbvh_free(&dv.legacy_bvh);
  dv.legacy_version = s.flat.version;
  const FlatScene& f = s.flat;
  const int ntris = (int)f.tris.size() / 3;
  if (ntris == 0) return 0;
  float *d_v = nullptr, *d_n = nullptr; int *d_t = nullptr, *d_m = nullptr;
  auto cleanup = [&]() { cudaFree(d_v); cudaFree(d_n); cudaFree(d_t); cudaFree(d_m); };
  cudaError_t e;
  if ((e = cudaMalloc(&d_v, f.verts.size() * 4)) != cudaSuccess || (e = cudaMalloc(&d_t, f.tris.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&d_n, f.n9.size() * 4)) != cudaSuccess || (e = cudaMalloc(&d_m, f.mat.size() * 4)) != cudaSuccess) { cleanup(); return fail_cuda(e, "cudaMalloc legacy scene"); }
  cudaMemcpyAsync(d_v, f.verts.data(), f.verts.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  cudaMemcpyAsync(d_t, f.tris.data(), f.tris.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  cudaMemcpyAsync(d_n, f.n9.data(), f.n9.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  cudaMemcpyAsync(d_m, f.mat.data(), f.mat.size() * 4, cudaMemcpyHostToDevice, dv.stream);
  e = gcore_dev::lbvh_build(d_v, (int)f.verts.size() / 3, d_t, ntris, d_n, d_m, &dv.legacy_bvh, dv.stream);
  cudaStreamSynchronize(dv.stream);
  cleanup();
  if (e != cudaSuccess) { gcore_dev::lbvh_free(&dv.legacy_bvh); return fail_cuda(e, "lbvh_build (legacy scene)"); }
  return 0;
}

void fill_view(art::BvhView& v, const LbvhBuffers& b) {
  v.nodes = b.nodes; v.tris = b.tris; v.tri_normals = b.tri_normals; v.tri_mat = b.tri_mat; v.ntris = b.ntris; v.root = b.root;
}

// ---- resident seam server.  Host side of the mailbox protocol (gcore_internal.h: SeamMailbox): `alive` goes 0 -> 1 here, before the
// launch, and 1 -> 0 in the kernel as its last store; a batch is posted by writing cmd and answered when done == its sequence number.
// The kernel may leave (idle) at the very moment a batch is posted: the poster sees alive == 0 and starts a new one, which finds the
// command waiting.  Everything here runs under State::mu.
static inline unsigned int mb_load(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void mb_store(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

static bool seam_server_enabled() {
  static const bool on = [] { const char* e = getenv("GCORE_SEAM_SERVER"); return !(e && e[0] == '0'); }();
  return on && seam().srv_usable;
}

// make the server leave and wait until it has (before the scene block in constant memory changes, before the device is released)
static void seam_server_stop() {
  Seam& q = seam();
  if (!q.h_mb || !q.srv_stream) return;
  if (mb_load(&q.h_mb->alive)) {
    mb_store(&q.h_mb->stop, 1u);
    cudaStreamSynchronize(q.srv_stream);
    mb_store(&q.h_mb->stop, 0u);
    mb_store(&q.h_mb->alive, 0u);                                       // (a kernel that died with an error never cleared it)
  }
}

static int seam_server_start(int two_sided) {
  State& s = S(); Seam& q = seam(); Device& dv = s.dev[0];
  if (!q.h_mb) {
    if (cudaHostAlloc(&q.h_mb, sizeof(gcore_dev::SeamMailbox), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&q.d_mb, q.h_mb, 0) != cudaSuccess) { q.h_mb = nullptr; return fail_cuda(cudaGetLastError(), "cudaHostAlloc seam mailbox"); }
    memset(q.h_mb, 0, sizeof(*q.h_mb));
    CU(cudaStreamCreateWithFlags(&q.srv_stream, cudaStreamNonBlocking), "cudaStreamCreate seam");
  }
  (void)dv;
  mb_store(&q.h_mb->stop, 0u);
  mb_store(&q.h_mb->alive, 1u);
  q.srv_two_sided = two_sided;
  const long long idle_ns = 2000000;                                    // 2 ms without a call: leave the SM (cudaMalloc / cudaFree wait for resident kernels)
  // `seen` = the last batch answered: a server started to replace one that left under a posted batch must find that batch new
  cudaError_t e = gcore_dev::launch_seam_server(q.d_mb, q.d_req[0], q.d_req[1], mb_load(&q.h_mb->done), idle_ns, two_sided, q.srv_stream);
  if (e != cudaSuccess) { mb_store(&q.h_mb->alive, 0u); return fail_cuda(e, "seam server launch"); }
  q.srv_starts++;
  return 0;
}

// post batch `buf` of n requests and wait for the answers; returns 1 when the server did not answer (caller falls back to a launch)
static int seam_server_batch(int buf, int n, int two_sided) {
  Seam& q = seam();
  if (q.h_mb && mb_load(&q.h_mb->alive) && q.srv_two_sided != two_sided) seam_server_stop();
  if (!q.h_mb || !mb_load(&q.h_mb->alive)) { if (seam_server_start(two_sided)) return 1; }
  q.seq = (q.seq + 1u) & 0xffffffu;
  const unsigned int seq = q.seq;
  mb_store(&q.h_mb->cmd, (seq << 8) | ((unsigned)buf << 7) | (unsigned)n);   // the requests were written before (release)
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned int spins = 1;; spins++) {
    if (mb_load(&q.h_mb->done) == seq) break;
    if ((spins & 0xfffu) == 0u) {
      if (!mb_load(&q.h_mb->alive)) {
        if (mb_load(&q.h_mb->done) == seq) break;                       // answered just before it left
        if (cudaStreamQuery(q.srv_stream) == cudaErrorNotReady) continue;   // `alive` cleared, kernel still retiring
        // left idle while the batch was being posted: a new one finds the command waiting.  If it cannot be started the batch goes to a
        // launch; nobody is resident, so `done` is ours to write: the posted command must not look new to a later server
        if (seam_server_start(two_sided)) { mb_store(&q.h_mb->done, seq); return 1; }
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) {   // never seen in practice; do not hang the caller on it
        seam_server_stop();
        mb_store(&q.h_mb->done, seq);
        q.srv_usable = false;
        return 1;
      }
    }
  }
  q.srv_batches++;
  return 0;
}

// make device d see the current scene: LBVHs of the right version + the constant block
int sync_scene(int d) {
  State& s = S(); Device& dv = s.dev[d];
  if (int rc = ensure_device(d)) return rc;
  // a resident seam kernel would keep reading the old constant block, and cudaMalloc / cudaFree below would wait for it to idle out
  if (d == 0 && (dv.scene_dirty || dv.mesh_version != s.mesh_version || dv.legacy_version != s.flat.version)) seam_server_stop();
  if (dv.mesh_version != s.mesh_version) { if (int rc = build_mesh_bvh(d)) return rc; dv.scene_dirty = true; }
  if (dv.legacy_version != s.flat.version) { if (int rc = build_legacy_bvh(d)) return rc; dv.scene_dirty = true; }
  if (dv.scene_dirty) {
    art::DeviceScene hs = s.hs;
    fill_view(hs.bvh_legacy, dv.legacy_bvh);
    if (s.use_committed) {                                       // the integrators trace the committed legacy scene
      fill_view(hs.bvh, dv.legacy_bvh);
      hs.ntris_bf = 0; hs.pair_group = 0; hs.mesh_mode = art::MESH_NEAREST;
      hs.mesh_kind = dv.legacy_bvh.ntris > 0 ? 2 : 0;
      for (int k = 0; k < 3; k++) { hs.mesh_tight_min[k] = s.flat.tight_min[k]; hs.mesh_tight_max[k] = s.flat.tight_max[k]; }
    } else {
      fill_view(hs.bvh, dv.mesh_bvh);
    }
    CU(gcore_dev::upload_scene(&hs, dv.stream), "upload_scene");
    CU(cudaStreamSynchronize(dv.stream), "upload_scene sync");   // `hs` is a stack copy
    dv.scene_dirty = false;
  }
  return 0;
}

int active_mesh_kind() {                                         // which render kernel family: 0/1 constant-memory mesh, 2 LBVH
  State& s = S();
  if (s.use_committed) return s.flat.tris.empty() ? 0 : 2;
  return s.hs.mesh_kind;
}

void free_accum(int d) {
  Device& dv = S().dev[d];
  if (dv.ready) {
    cudaSetDevice(dv.ordinal);
    if (dv.d_gather) cudaFree(dv.d_gather);
    else if (dv.d_accum) cudaFree(dv.d_accum);
  }
  dv.d_accum = dv.d_gather = nullptr;
  dv.nslots = 0;
}

int alloc_accum() {
  State& s = S();
  const int tiles_x = (s.width + GC_TILE - 1) / GC_TILE, tiles_y = (s.height + GC_TILE - 1) / GC_TILE;
  const int ntiles = tiles_x * tiles_y;
  const int we = world_eff();
  const int local_tiles = (ntiles + we - 1) / we;              // equal share per rank (padded) so gathers are uniform
  const int nslots = local_tiles * GC_TILE * GC_TILE;
  for (int d = 0; d < s.gpu_count; d++) {
    if (int rc = ensure_device(d)) return rc;
    Device& dv = s.dev[d];
    if (s.width <= 0 || s.height <= 0) { free_accum(d); continue; }
    if (!dv.d_accum || nslots != dv.nslots) {                   // same size: keep the allocation, just clear it
      free_accum(d);
      if (d == 0) {                                             // device 0 receives the other devices' accumulators behind its own
        CU(cudaMalloc(&dv.d_gather, sizeof(float4) * (size_t)nslots * s.gpu_count), "cudaMalloc accum");
        dv.d_accum = dv.d_gather;
      } else CU(cudaMalloc(&dv.d_accum, sizeof(float4) * (size_t)nslots), "cudaMalloc accum");
      dv.nslots = nslots;
    }
    CU(cudaMemsetAsync(dv.d_accum, 0, sizeof(float4) * (size_t)dv.nslots, dv.stream), "memset accum");
    CU(cudaMemsetAsync(dv.d_rays, 0, 6 * sizeof(unsigned long long), dv.stream), "memset counters");
    memset(dv.last_counters, 0, sizeof(dv.last_counters));
  }
  s.spp = 0; s.palette_frame = false;
  return 0;
}

// work buffers of render_wave_kernel: per-slot recursion stacks (fixed size) and the (pass, pixel) colour buffer.
// The colour buffer is sized so that every BASELINE config is one launch (C4: 2048^2 x 64 units x 16 B = 4.3 GB of the
// 180 GB of HBM3e); longer renders go in chunks of passes.
#define GC_COLBUF_CAP_BYTES ((size_t)8 << 30)
int ensure_wave_buffers(int d, size_t items) {
  Device& dv = S().dev[d];
  if (!dv.wave.sm_count) {
    const size_t total_slots = (size_t)gcore_dev::wave_total_slots(dv.sm_count);
    float4* a = nullptr; float2* b = nullptr; unsigned int* c = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&a, sizeof(float4) * total_slots * gcore_dev::wave_max_depth())) != cudaSuccess ||
        (e = cudaMalloc(&b, sizeof(float2) * total_slots * gcore_dev::wave_max_depth())) != cudaSuccess ||
        (e = cudaMalloc(&c, sizeof(unsigned int))) != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return fail_cuda(e, "cudaMalloc wave stacks");
    }
    // the tail of a launch: twice the resident path slots' worth of items go out one sub-pixel path at a time (render_wave.cuh wave_regen)
    int tail_items = 2 * (int)total_slots;
    if (const char* e = getenv("GCORE_TAIL_ITEMS")) tail_items = atoi(e) < 0 ? 0 : atoi(e);               // experiments: 0 switches the split off
    float4* t = nullptr;
    if ((e = cudaMalloc(&t, sizeof(float4) * 4 * (size_t)(tail_items > 0 ? tail_items : 1))) != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return fail_cuda(e, "cudaMalloc wave tail");
    }
    dv.wave.tailbuf = t; dv.wave.tail_items = tail_items;
    dv.wave.stackA = a; dv.wave.stackB = b; dv.wave.next_item = c;
    dv.wave.sm_count = dv.sm_count;                             // only now: a failed allocation leaves the buffers unset
  }
  if (items > dv.wave.colbuf_items) {
    if (dv.wave.colbuf) { cudaFree(dv.wave.colbuf); dv.wave.colbuf = nullptr; dv.wave.colbuf_items = 0; }
    CU(cudaMalloc(&dv.wave.colbuf, sizeof(float4) * items), "cudaMalloc colbuf");
    dv.wave.colbuf_items = items;
  }
  return 0;
}

RenderArgs make_args(int d) {
  State& s = S(); Device& dv = s.dev[d];
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.width = s.width; a.height = s.height;
  a.tiles_x = (s.width + GC_TILE - 1) / GC_TILE;
  a.ntiles = a.tiles_x * ((s.height + GC_TILE - 1) / GC_TILE);
  a.rank = rank_eff(d); a.world = world_eff(); a.nslots = dv.nslots;
  a.anti_aliasing = s.cfg.anti_aliasing ? 1 : 0; a.max_depth = s.cfg.max_depth; a.passes = 0;
  a.seed = s.cfg.seed; a.first_sample = s.spp;
  a.bg[0] = s.cfg.background[0]; a.bg[1] = s.cfg.background[1]; a.bg[2] = s.cfg.background[2];
  const float fov = 1.57079632679489661923f;                      // Pi / 2.0            (ray_tracer.adb:63)
  a.zdir = -(float)s.width / tanf(fov / 2.0f);                     // safe_tan(fov / 2.0) (ray_tracer.adb:67, vector_math.adb:14-22)
  a.half_w = (float)s.width / 2.0f; a.half_h = (float)s.height / 2.0f;
  a.accum = dv.d_accum; a.counters = dv.d_rays;
  return a;
}

int ensure_out_buffers(bool hdr, bool ldr) {                     // device 0
  State& s = S(); Device& dv = s.dev[0];
  size_t px = (size_t)s.width * s.height;
  if (px != dv.out_pixels) {
    if (dv.d_hdr) { cudaFree(dv.d_hdr); dv.d_hdr = nullptr; }
    if (dv.d_ldr) { cudaFree(dv.d_ldr); dv.d_ldr = nullptr; }
    dv.out_pixels = px;
  }
  if (hdr && !dv.d_hdr) CU(cudaMalloc(&dv.d_hdr, sizeof(float4) * px), "cudaMalloc hdr");
  if (ldr && !dv.d_ldr) CU(cudaMalloc(&dv.d_ldr, sizeof(uint32_t) * px), "cudaMalloc ldr");
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

// padded tight bounds of a vertex set (culling only; 1e-4 relative + 1e-5 absolute >> rounding error of the triangle test)
void tight_bounds(const std::vector<float>& pos, float mn_o[3], float mx_o[3]) {
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (size_t i = 0; i + 2 < pos.size(); i += 3)
    for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], pos[i + k]); mx[k] = fmaxf(mx[k], pos[i + k]); }
  for (int k = 0; k < 3; k++) {
    float pad = fmaxf(fabsf(mn[k]), fabsf(mx[k])) * 1.0e-4f + 1.0e-5f;
    mn_o[k] = mn[k] - pad; mx_o[k] = mx[k] + pad;
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

// the shaded mesh of gcore_set_mesh / gcore_set_mesh_vsgf: constant-memory triangles up to 64, LBVH per device above
int install_mesh(int mesh_mode) {
  State& s = S();
  const int ntris = (int)s.mesh_tris.size() / 3;
  const int nverts = (int)s.mesh_pos.size() / 3;
  bool bad_index = false;
  for (int v : s.mesh_tris) if (v < 0 || v >= nverts) bad_index = true;
  const bool too_big = ntris > GC_MAX_BF_TRIS && mesh_mode == GCORE_MESH_REFERENCE_BF;
  if (bad_index || too_big) { s.mesh_pos.clear(); s.mesh_nrm.clear(); s.mesh_tris.clear(); }   // leave a scene without a mesh, not half of one
  s.mesh_version++;
  s.hs.mesh_mode = mesh_mode;
  s.hs.ntris_bf = 0; s.hs.mesh_kind = 0; s.hs.pair_group = 0;
  for (int k = 0; k < 3; k++) { s.hs.mesh_bmin[k] = s.mesh_bbox[k]; s.hs.mesh_bmax[k] = s.mesh_bbox[3 + k]; }
  mark_scene_dirty();
  if (bad_index) return fail(GCORE_E_ARG, "gcore_set_mesh: triangle index out of range");
  if (too_big) return fail(GCORE_E_LIMIT, "GCORE_MESH_REFERENCE_BF (brute-force reference semantics) is limited to 64 triangles; use GCORE_MESH_NEAREST");
  if (ntris == 0) return 0;
  tight_bounds(s.mesh_pos, s.hs.mesh_tight_min, s.hs.mesh_tight_max);
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
  s.hs.mesh_kind = 2;
  // build on the devices in use right away: an allocation / build failure belongs to this call
  for (int d = 0; d < s.gpu_count; d++) { if (int rc = ensure_device(d)) return rc; if (int rc = build_mesh_bvh(d)) return rc; s.dev[d].scene_dirty = true; }
  return 0;
}

// normC = 1 / g_spp for path-traced frames; the debug palette is resolved as it is (ray_tracer.adb:245-257 vs :281-291)
int resolve_spp() {
  State& s = S();
  return (s.palette_frame || s.spp <= 0) ? 1 : s.spp;
}

int render_units_locked(int units, gcore_render_stats* stats) {
  State& s = S();
  if (units < 0) return fail(GCORE_E_ARG, "units < 0");
  if (!s.dev[0].d_accum) return fail(GCORE_E_STATE, "gcore_resize_viewport has not been called");
  if (s.cfg.max_depth < 1 || s.cfg.max_depth > 16) return fail(GCORE_E_ARG, "max_depth must be in 1..16");
  const int G = s.gpu_count;
  seam_server_stop();                                            // the render kernels want every SM: a resident seam kernel leaves now, not 2 ms from now
  for (int d = 0; d < G; d++) if (int rc = sync_scene(d)) return rc;
  const int mesh_kind = active_mesh_kind();
  int launches = 0;
  uint64_t samples = 0;
  const uint64_t pixels = (uint64_t)s.width * s.height;
  const bool debug = (s.cfg.render_type == GCORE_RT_DEBUG || s.cfg.render_type == GCORE_RT_WHITTED);
  const int per = s.cfg.anti_aliasing ? 4 : 1;
  size_t chunk = 0;
  if (!debug && units > 0) {
    // the colour buffer holds one float4 per (pass, pixel): very long renders go in chunks of passes
    size_t cap_bytes = GC_COLBUF_CAP_BYTES;
    if (const char* e = getenv("GCORE_COLBUF_CAP_KB")) { long v = atol(e); if (v > 0) cap_bytes = (size_t)v << 10; }   // tests: force the chunked path on small frames
    const size_t nslots = (size_t)s.dev[0].nslots;
    chunk = cap_bytes / (sizeof(float4) * nslots);
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)units) chunk = (size_t)units;
    while (chunk > 1 && ((chunk * nslots) >> 30)) chunk >>= 1;                 // (the counter also runs over the tail items' sub-paths)
    if ((chunk * nslots) >> 30) return fail(GCORE_E_LIMIT, "viewport too large for the work-item counter");
  }
  // every device gets its whole share enqueued before the next one: launches are asynchronous, the devices run concurrently
  for (int d = 0; d < G; d++) {
    Device& dv = s.dev[d];
    CU(cudaSetDevice(dv.ordinal), "cudaSetDevice");
    RenderArgs a = make_args(d);
    a.passes = units;
    CU(cudaEventRecord(dv.ev0, dv.stream), "event");
    if (debug) {
      CU(gcore_dev::launch_debug(mesh_kind, a, dv.stream), "debug_kernel");                 // Debug_Ray_Tracing overwrites the buffer
      launches += 1;
    } else if (units > 0) {
      if (int rc = ensure_wave_buffers(d, chunk * (size_t)dv.nslots)) return rc;
      int spp = s.spp;
      for (int done = 0; done < units; done += (int)chunk) {
        const int n = (units - done < (int)chunk) ? (units - done) : (int)chunk;
        a.passes = n; a.first_sample = spp;
        CU(gcore_dev::launch_render_wave(s.cfg.render_type, mesh_kind, a, dv.wave, n, s.stats_level >= 2 ? 1 : 0, dv.stream), "render_wave_kernel");
        launches += 2;
        spp += n * per;
      }
    }
    CU(cudaEventRecord(dv.ev1, dv.stream), "event");
  }
  if (debug) { samples = pixels; s.palette_frame = true; }                                    // g_spp is not touched (ray_tracer.adb:245-260)
  else if (units > 0) { samples = pixels * (uint64_t)(units * per); s.spp += units * per; s.palette_frame = false; }   // ray_tracer.adb:168-176
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    float kernel_ms = 0.f;
    for (int d = 0; d < G; d++) {
      Device& dv = s.dev[d];
      CU(cudaSetDevice(dv.ordinal), "cudaSetDevice");
      unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
      CU(cudaMemcpyAsync(c, dv.d_rays, sizeof(c), cudaMemcpyDeviceToHost, dv.stream), "memcpy counters");
      CU(cudaStreamSynchronize(dv.stream), "render sync");
      float ms = 0.f; cudaEventElapsedTime(&ms, dv.ev0, dv.ev1);
      dv.last_kernel_ms = ms;
      if (ms > kernel_ms) kernel_ms = ms;
      stats->rays += c[0] - dv.last_counters[0]; stats->mesh_queries += c[1] - dv.last_counters[1]; stats->bounces += c[2] - dv.last_counters[2];
      for (int k = 0; k < 6; k++) dv.last_counters[k] = c[k];
    }
    stats->samples = samples; stats->kernel_ms = kernel_ms; stats->total_ms = kernel_ms; stats->spp = s.spp;
    stats->kernel_launches = launches;
  }
  return 0;
}

// ---- the one exchange step of a multi-GPU frame: accumulators of devices 1..G-1 -> device 0, behind its own (NCCL)
int ensure_nccl() {
  State& s = S();
  if (s.nccl_ready) return 0;
  NcclApi& n = s.nccl;
  if (!n.h) {
    n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) return fail(GCORE_E_STATE, std::string("multi-GPU rendering needs NCCL: ") + dlerror());
    n.CommInitAll = (decltype(n.CommInitAll))dlsym(n.h, "ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.h, "ncclCommDestroy");
    n.GroupStart = (decltype(n.GroupStart))dlsym(n.h, "ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.h, "ncclGroupEnd");
    n.Send = (decltype(n.Send))dlsym(n.h, "ncclSend");
    n.Recv = (decltype(n.Recv))dlsym(n.h, "ncclRecv");
    n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.h, "ncclGetErrorString");
    if (!n.CommInitAll || !n.CommDestroy || !n.GroupStart || !n.GroupEnd || !n.Send || !n.Recv || !n.GetErrorString)
      return fail(GCORE_E_STATE, "libnccl lacks an entry point");
  }
  int devs[GC_MAX_GPUS];
  for (int d = 0; d < s.gpu_count; d++) devs[d] = s.dev[d].ordinal;
  ncclResult_t r = n.CommInitAll(s.comms, s.gpu_count, devs);
  if (r != ncclSuccess) return fail(GCORE_E_CUDA, std::string("ncclCommInitAll: ") + n.GetErrorString(r));
  s.nccl_ready = true;
  return 0;
}

void destroy_nccl() {
  State& s = S();
  if (!s.nccl_ready) return;
  for (int d = 0; d < s.gpu_count; d++) s.nccl.CommDestroy(s.comms[d]);
  s.nccl_ready = false;
}

int gather_accumulators() {
  State& s = S();
  const int G = s.gpu_count;
  if (G == 1) return 0;
  if (int rc = ensure_nccl()) return rc;
  NcclApi& n = s.nccl;
  const size_t count = (size_t)s.dev[0].nslots * 4;
  ncclResult_t r = n.GroupStart();
  for (int d = 1; d < G && r == ncclSuccess; d++) {
    r = n.Send(s.dev[d].d_accum, count, ncclFloat, 0, s.comms[d], s.dev[d].stream);
    if (r == ncclSuccess) r = n.Recv(s.dev[0].d_gather + (size_t)d * s.dev[0].nslots, count, ncclFloat, d, s.comms[0], s.dev[0].stream);
  }
  ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return fail(GCORE_E_CUDA, std::string("nccl gather: ") + n.GetErrorString(r));
  return 0;
}

// resolve into device buffers of device 0 (gathering first when several devices rendered)
int resolve_to(float4* d_hdr, uint32_t* d_ldr, int* launches) {
  State& s = S(); Device& d0 = s.dev[0];
  if (int rc = gather_accumulators()) return rc;
  CU(cudaSetDevice(d0.ordinal), "cudaSetDevice");
  RenderArgs a = make_args(0);
  if (s.gpu_count == 1) CU(gcore_dev::launch_resolve(a, resolve_spp(), d_hdr, d_ldr, d0.stream), "resolve_kernel");
  else CU(gcore_dev::launch_resolve_gathered(a, d0.d_gather, resolve_spp(), d_hdr, d_ldr, d0.stream), "resolve_kernel");
  if (launches) *launches += 1;
  return 0;
}

int read_frame_locked(float* hdr4, uint32_t* ldr, int* launches) {
  State& s = S(); Device& d0 = s.dev[0];
  if (!d0.d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (s.world != 1) return fail(GCORE_E_STATE, "gcore_read_frame needs world == 1 (multi-process partition: gather the accumulators and use gcore_resolve_gathered_device; one process: gcore_set_gpu_count)");
  if (!hdr4 && !ldr) return 0;
  CU(cudaSetDevice(d0.ordinal), "cudaSetDevice");
  if (int rc = ensure_out_buffers(hdr4 != nullptr, ldr != nullptr)) return rc;
  if (int rc = resolve_to(hdr4 ? d0.d_hdr : nullptr, ldr ? d0.d_ldr : nullptr, launches)) return rc;
  size_t px = (size_t)s.width * s.height;
  if (hdr4) CU(cudaMemcpyAsync(hdr4, d0.d_hdr, sizeof(float4) * px, cudaMemcpyDeviceToHost, d0.stream), "memcpy hdr");
  if (ldr) CU(cudaMemcpyAsync(ldr, d0.d_ldr, sizeof(uint32_t) * px, cudaMemcpyDeviceToHost, d0.stream), "memcpy ldr");
  CU(cudaStreamSynchronize(d0.stream), "read_frame sync");
  return 0;
}

void release_device(int d) {
  State& s = S(); Device& dv = s.dev[d];
  if (!dv.ready) return;
  if (d == 0) seam_server_stop();
  cudaSetDevice(dv.ordinal);
  cudaStreamSynchronize(dv.stream);
  free_accum(d);
  gcore_dev::lbvh_free(&dv.mesh_bvh); gcore_dev::lbvh_free(&dv.legacy_bvh);
  dv.mesh_version = dv.legacy_version = -1;
  cudaFree(dv.wave.colbuf); cudaFree(dv.wave.stackA); cudaFree(dv.wave.stackB); cudaFree(dv.wave.next_item); cudaFree(dv.wave.tailbuf);
  dv.wave = WaveBuffers();
  dv.scene_dirty = true;
}

// flatten the instances of the legacy scene into world space (what rtcCommitScene of the instanced scene makes traceable,
// cpp/embree_connect.cpp:147-184, 241-244)
void flatten_legacy() {
  State& s = S();
  FlatScene f;
  f.version = s.flat.version + 1;
  std::vector<LegacyInstance> inst = s.instances;
  if (inst.empty()) {                                                  // no instance declared: every mesh once, identity
    for (int m = 0; m < (int)s.meshes.size(); m++) { LegacyInstance in; in.mesh = m; memset(in.m, 0, sizeof(in.m)); in.m[0] = in.m[5] = in.m[10] = 1.0f; inst.push_back(in); }
  }
  for (int ii = 0; ii < (int)inst.size(); ii++) {
    const LegacyMesh& m = s.meshes[inst[ii].mesh];
    const float* M = inst[ii].m;
    const int base = (int)f.verts.size() / 3;
    for (size_t v = 0; v + 2 < m.verts.size(); v += 3) {
      float x = m.verts[v], y = m.verts[v + 1], z = m.verts[v + 2];
      f.verts.push_back(M[0] * x + M[1] * y + M[2] * z + M[3]);
      f.verts.push_back(M[4] * x + M[5] * y + M[6] * z + M[7]);
      f.verts.push_back(M[8] * x + M[9] * y + M[10] * z + M[11]);
    }
    const bool has_n = !m.normals.empty() && m.normals.size() == m.verts.size();
    for (size_t t = 0; t + 2 < m.idx.size(); t += 3) {
      const int i3[3] = {m.idx[t], m.idx[t + 1], m.idx[t + 2]};
      for (int c = 0; c < 3; c++) f.tris.push_back(base + i3[c]);
      f.prim.push_back((int)(t / 3)); f.inst.push_back(ii); f.mat.push_back(m.mat);
      float n9[9];
      if (has_n) {                                                     // vertex normals under the instance's 3x3, normalised
        for (int c = 0; c < 3; c++) {
          const float nx = m.normals[3 * (size_t)i3[c]], ny = m.normals[3 * (size_t)i3[c] + 1], nz = m.normals[3 * (size_t)i3[c] + 2];
          const float wx = M[0] * nx + M[1] * ny + M[2] * nz, wy = M[4] * nx + M[5] * ny + M[6] * nz, wz = M[8] * nx + M[9] * ny + M[10] * nz;
          const float l_inv = 1.0f / sqrtf(wx * wx + wy * wy + wz * wz);
          n9[3 * c] = l_inv * wx; n9[3 * c + 1] = l_inv * wy; n9[3 * c + 2] = l_inv * wz;
        }
      } else {                                                         // geometric normal normalize(cross(B - A, C - A)) at all three corners
        const float* A = &f.verts[3 * (size_t)(base + i3[0])]; const float* B = &f.verts[3 * (size_t)(base + i3[1])]; const float* C = &f.verts[3 * (size_t)(base + i3[2])];
        const float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        const float cx = e1[1] * e2[2] - e2[1] * e1[2], cy = e1[2] * e2[0] - e2[2] * e1[0], cz = e1[0] * e2[1] - e2[0] * e1[1];
        const float l_inv = 1.0f / sqrtf(cx * cx + cy * cy + cz * cz);
        for (int c = 0; c < 3; c++) { n9[3 * c] = l_inv * cx; n9[3 * c + 1] = l_inv * cy; n9[3 * c + 2] = l_inv * cz; }
      }
      f.n9.insert(f.n9.end(), n9, n9 + 9);
    }
  }
  if (!f.verts.empty()) tight_bounds(f.verts, f.tight_min, f.tight_max);
  s.flat = std::move(f);
  mark_scene_dirty();
}

// device staging of the batched seam queries: grown on demand and kept
int ensure_batch_buffers(size_t n) {
  Device& dv = S().dev[0];
  if (n <= dv.batch_cap) return 0;
  cudaFree(dv.d_batch_rays); cudaFree(dv.d_batch_hits); cudaFree(dv.d_batch_found);
  dv.d_batch_rays = nullptr; dv.d_batch_hits = nullptr; dv.d_batch_found = nullptr; dv.batch_cap = 0;
  size_t cap = n < 256 ? 256 : n;
  CU(cudaMalloc(&dv.d_batch_rays, sizeof(gcore_ray) * cap), "cudaMalloc batch");
  CU(cudaMalloc(&dv.d_batch_hits, sizeof(gcore_HitCpp) * cap), "cudaMalloc batch");
  CU(cudaMalloc(&dv.d_batch_found, sizeof(int) * cap), "cudaMalloc batch");
  dv.batch_cap = cap;
  return 0;
}

int read_counters(unsigned long long out[6]) {
  State& s = S();
  for (int k = 0; k < 6; k++) out[k] = 0;
  for (int d = 0; d < s.gpu_count; d++) {
    if (int rc = ensure_device(d)) return rc;
    Device& dv = s.dev[d];
    unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(c, dv.d_rays, sizeof(c), cudaMemcpyDeviceToHost, dv.stream), "memcpy counters");
    CU(cudaStreamSynchronize(dv.stream), "counters sync");
    for (int k = 0; k < 6; k++) out[k] += c[k];
  }
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
  if (device < 0) return fail(GCORE_E_ARG, "gcore_set_device: negative device");
  if (s.dev[0].ready && device != s.base_device) return fail(GCORE_E_STATE, "gcore_set_device must be called before any other device work");
  s.base_device = device;
  return ensure_device(0);
}

int gcore_set_gpu_count(int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (n < 1 || n > GC_MAX_GPUS) return fail(GCORE_E_ARG, "gcore_set_gpu_count: 1..8 devices");
  if (n > 1 && s.world != 1) return fail(GCORE_E_STATE, "gcore_set_gpu_count: in-process multi-GPU and the multi-process partition (gcore_set_partition) exclude each other");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { cudaGetLastError(); return fail(GCORE_E_CUDA, "no CUDA device available (libgcore_b200 has no CPU fallback)"); }
  if (s.base_device + n > have) return fail(GCORE_E_ARG, "gcore_set_gpu_count: more devices than the box has");
  if (n == s.gpu_count) return 0;
  destroy_nccl();
  for (int d = 0; d < GC_MAX_GPUS; d++) { if (d >= n) release_device(d); else free_accum(d); }
  s.gpu_count = n;
  s.spp = 0; s.palette_frame = false;
  if (s.width > 0) return alloc_accum();
  return 0;
  GC_CATCH_INT
}

int gcore_get_gpu_count(void) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  return s.gpu_count;
}

int gcore_set_partition(int rank, int world) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (world < 1 || rank < 0 || rank >= world) return fail(GCORE_E_ARG, "bad partition");
  if (world > 1 && s.gpu_count != 1) return fail(GCORE_E_STATE, "gcore_set_partition: the multi-process partition and in-process multi-GPU (gcore_set_gpu_count) exclude each other");
  s.rank = rank; s.world = world;
  if (s.width > 0) return alloc_accum();
  return 0;
}

int gcore_stream(void** stream) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = ensure_device(0)) return rc;
  if (stream) *stream = (void*)s.dev[0].stream;
  return 0;
}

int gcore_synchronize(void) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  for (int d = 0; d < s.gpu_count; d++) {
    if (int rc = ensure_device(d)) return rc;
    CU(cudaStreamSynchronize(s.dev[d].stream), "synchronize");
  }
  return 0;
}

int gcore_bvh_info(int64_t* ntris, int64_t* nnodes, int64_t* bytes, float* build_ms) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  const LbvhBuffers& b = (!s.use_committed && s.dev[0].mesh_bvh.ntris) ? s.dev[0].mesh_bvh : s.dev[0].legacy_bvh;
  if (ntris) *ntris = b.ntris;
  if (nnodes) *nnodes = b.nnodes;
  if (bytes) *bytes = (int64_t)b.bytes;
  if (build_ms) *build_ms = b.build_ms;
  return 0;
}

int gcore_kernel_regs(int render_type) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = ensure_device(0)) return rc;
  return gcore_dev::render_kernel_regs(render_type);
}

int gcore_device_kernel_ms(float* out, int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!out || n < 0) return fail(GCORE_E_ARG, "gcore_device_kernel_ms: bad arguments");
  for (int d = 0; d < n; d++) out[d] = d < s.gpu_count ? s.dev[d].last_kernel_ms : 0.f;
  return 0;
}

// ---------------------------------------------------------------------------------------------- legacy seam
void gcore_destroy(void) {                                             // embree_connect.cpp:51-58
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  s.meshes.clear(); s.instances.clear();
  const int v = s.flat.version + 1;
  s.flat = FlatScene(); s.flat.version = v;
  for (int d = 0; d < GC_MAX_GPUS; d++) if (s.dev[d].ready) { cudaSetDevice(s.dev[d].ordinal); gcore_dev::lbvh_free(&s.dev[d].legacy_bvh); s.dev[d].legacy_version = v; }
  mark_scene_dirty();
  s.legacy_alive = false;
  GC_CATCH_VOID
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
  try {
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
  } catch (...) { fail(GCORE_E_LIMIT, "gcore_add_mesh_3f: out of host memory"); return -1; }
}

int gcore_set_mesh_material(int meshId, int matId) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (meshId < 0 || meshId >= (int)s.meshes.size()) return fail(GCORE_E_ARG, "gcore_set_mesh_material: bad meshId");
  if (matId < 0 || matId >= GC_MAX_MATERIALS) return fail(GCORE_E_ARG, "gcore_set_mesh_material: material id outside 0..31");
  s.meshes[meshId].mat = matId;
  return 0;
}

int gcore_set_mesh_normals_3f(int meshId, const float* a_normals3f, int a_vertexNum) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (meshId < 0 || meshId >= (int)s.meshes.size()) return fail(GCORE_E_ARG, "gcore_set_mesh_normals_3f: bad meshId");
  LegacyMesh& m = s.meshes[meshId];
  if (!a_normals3f) { m.normals.clear(); return 0; }
  if (a_vertexNum < 0 || (size_t)a_vertexNum * 3 < m.verts.size()) return fail(GCORE_E_ARG, "gcore_set_mesh_normals_3f: fewer normals than vertices in use");
  m.normals.assign(a_normals3f, a_normals3f + m.verts.size());
  return 0;
  GC_CATCH_INT
}

void gcore_instance_meshes(int a_geomId, const float* a_matrices16f, int a_matrixNum) {   // :147-184
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  s.err_code = 0; s.err_text.clear();
  if (a_geomId < 0 || a_geomId >= (int)s.meshes.size() || !a_matrices16f) { fail(GCORE_E_ARG, "gcore_instance_meshes: bad meshId"); return; }
  for (int k = 0; k < a_matrixNum; k++) {
    LegacyInstance in; in.mesh = a_geomId;
    memcpy(in.m, a_matrices16f + 16 * (size_t)k, 12 * sizeof(float));   // RTC_FORMAT_FLOAT3X4_ROW_MAJOR (:169)
    s.instances.push_back(in);
  }
  GC_CATCH_VOID
}

void gcore_commit_scene(void) {                                        // :241-244
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  s.err_code = 0; s.err_text.clear();                                  // void signature: gcore_last_error() afterwards speaks for this call
  if (ensure_device(0)) return;
  flatten_legacy();
  for (int d = 0; d < s.gpu_count; d++) { if (ensure_device(d)) return; if (build_legacy_bvh(d)) return; s.dev[d].scene_dirty = true; }
  GC_CATCH_VOID
}

int gcore_use_committed_scene(int enable) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  s.use_committed = enable != 0;
  mark_scene_dirty();
  return 0;
}

int gcore_set_seam_two_sided(int enable) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  s.seam_two_sided = enable ? 1 : 0;
  return 0;
}

int gcore_bvh_closest_hit_batch(int n, const gcore_ray* rays, float t_near, float t_far, gcore_HitCpp* hits, int32_t* found) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (n < 0 || (n > 0 && (!rays || !hits || !found))) return fail(GCORE_E_ARG, "gcore_bvh_closest_hit_batch: bad arguments");
  if (n == 0) return 0;
  if (int rc = sync_scene(0)) return rc;
  if (int rc = ensure_batch_buffers((size_t)n)) return rc;
  Device& dv = s.dev[0];
  gcore_ray* d_r = dv.d_batch_rays; gcore_HitCpp* d_h = dv.d_batch_hits; int* d_f = dv.d_batch_found;
  cudaError_t e = cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_h, hits, sizeof(gcore_HitCpp) * (size_t)n, cudaMemcpyHostToDevice, dv.stream);   // records are only written on a hit
  if (e == cudaSuccess) e = gcore_dev::launch_bvh_batch(n, d_r, t_near, t_far, d_h, d_f, s.seam_two_sided, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_HitCpp) * (size_t)n, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(found, d_f, 4 * (size_t)n, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
  if (e != cudaSuccess) return fail_cuda(e, "bvh batch");
  // flattened triangle id -> (primIndex inside its mesh, instIndex); Ng from the flattened world-space triangle
  for (int i = 0; i < n; i++) if (found[i]) {
    int flat = hits[i].primIndex;
    if (flat >= 0 && flat < (int)s.flat.prim.size()) { hits[i].primIndex = s.flat.prim[flat]; hits[i].instIndex = s.flat.inst[flat]; hits[i].geomIndex = 0; }
  }
  return 0;
  GC_CATCH_INT
}

int gcore_bvh_occluded_batch(int n, const gcore_ray* rays, float t_near, float t_far, int32_t* occluded) {   // rtcOccluded1 counterpart
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (n < 0 || (n > 0 && (!rays || !occluded))) return fail(GCORE_E_ARG, "gcore_bvh_occluded_batch: bad arguments");
  if (n == 0) return 0;
  if (int rc = sync_scene(0)) return rc;
  if (int rc = ensure_batch_buffers((size_t)n)) return rc;
  Device& dv = s.dev[0];
  cudaError_t e = cudaMemcpyAsync(dv.d_batch_rays, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, dv.stream);
  if (e == cudaSuccess) e = gcore_dev::launch_bvh_occluded(n, dv.d_batch_rays, t_near, t_far, dv.d_batch_found, s.seam_two_sided, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(occluded, dv.d_batch_found, 4 * (size_t)n, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
  if (e != cudaSuccess) return fail_cuda(e, "bvh occluded");
  return 0;
  GC_CATCH_INT
}

int gcore_bvh_bench(int n, const gcore_ray* rays, float t_near, float t_far, int which, int repeat, float* ms_per_launch, uint64_t* stats3) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (n <= 0 || !rays || repeat < 1 || which < 0 || which > 1) return fail(GCORE_E_ARG, "gcore_bvh_bench: bad arguments");
  if (int rc = sync_scene(0)) return rc;
  Device& dv = s.dev[0];
  const LbvhBuffers& tree = which ? (s.use_committed ? dv.legacy_bvh : dv.mesh_bvh) : dv.legacy_bvh;
  if (tree.ntris <= 0 || (tree.ntris > 1 && !tree.nodes)) return fail(GCORE_E_STATE, "gcore_bvh_bench: no LBVH of that kind is built");
  gcore_ray* d_r = nullptr; gcore_HitCpp* d_h = nullptr; int* d_f = nullptr; unsigned long long* d_c = nullptr;
  auto cleanup = [&]() { cudaFree(d_r); cudaFree(d_h); cudaFree(d_f); cudaFree(d_c); };
  cudaError_t e;
  if ((e = cudaMalloc(&d_r, sizeof(gcore_ray) * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_h, sizeof(gcore_HitCpp) * (size_t)n)) != cudaSuccess ||
      (e = cudaMalloc(&d_f, 4 * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_c, 24)) != cudaSuccess) { cleanup(); return fail_cuda(e, "cudaMalloc bvh bench"); }
  cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, dv.stream);
  cudaMemsetAsync(d_c, 0, 24, dv.stream);
  e = gcore_dev::launch_bvh_batch_which(n, d_r, t_near, t_far, d_h, d_f, which, d_c, dv.stream);          // counting pass (also the warm-up)
  unsigned long long c[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(c, d_c, 24, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaEventRecord(dv.ev0, dv.stream);
  for (int k = 0; k < repeat && e == cudaSuccess; k++) e = gcore_dev::launch_bvh_batch_which(n, d_r, t_near, t_far, d_h, d_f, which, nullptr, dv.stream);
  if (e == cudaSuccess) e = cudaEventRecord(dv.ev1, dv.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
  cleanup();
  if (e != cudaSuccess) return fail_cuda(e, "bvh bench");
  float ms = 0.f; cudaEventElapsedTime(&ms, dv.ev0, dv.ev1);
  if (ms_per_launch) *ms_per_launch = ms / (float)repeat;
  if (stats3) { stats3[0] = c[0]; stats3[1] = c[1]; stats3[2] = c[2]; }
  return 0;
  GC_CATCH_INT
}

// one combined batch: called by the leader without the seam mutex
static int seam_launch(int buf, int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (int rc = sync_scene(0)) return rc;
  Device& dv = s.dev[0];
  if (dv.legacy_bvh.ntris <= 0) { for (int i = 0; i < n; i++) seam().h_req[buf]->out[i].found = 0; return 0; }
  if (!(seam_server_enabled() && seam_server_batch(buf, n, s.seam_two_sided) == 0)) {
    cudaError_t e = gcore_dev::launch_bvh_seam(n, seam().d_req[buf], s.seam_two_sided, dv.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
    if (e != cudaSuccess) return fail_cuda(e, "gcore_closest_hit");
  }
  // flattened triangle id -> (primIndex inside its mesh, instIndex), here, under the scene lock the leader already holds: a caller
  // that took the lock for its own hit would queue behind the next batch in flight
  for (int i = 0; i < n; i++) {
    gcore_dev::SeamOut& o = seam().h_req[buf]->out[i];
    const int flat = o.hit.primIndex;
    if (o.found && flat >= 0 && flat < (int)s.flat.prim.size()) { o.hit.primIndex = s.flat.prim[flat]; o.hit.instIndex = s.flat.inst[flat]; o.hit.geomIndex = 0; }
  }
  return 0;
}

bool gcore_closest_hit(const float a_rayPos[3], const float a_rayDir[3], float t_near, float t_far, gcore_HitCpp* pHit) {   // :196-238
  if (!a_rayPos || !a_rayDir || !pHit) return false;
  try {
    Seam& q = seam();
    std::unique_lock<SpinLock> lk(q.m);
    if (!q.h_req[0]) {                                                  // first call: the request / answer buffers
      State& s = S();
      std::lock_guard<std::recursive_mutex> slk(s.mu);
      if (ensure_device(0)) return false;
      for (int b = 0; b < 2; b++) {
        if (cudaHostAlloc(&q.h_req[b], sizeof(gcore_dev::SeamBatch), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&q.d_req[b], q.h_req[b], 0) != cudaSuccess) { fail_cuda(cudaGetLastError(), "cudaHostAlloc seam"); q.h_req[0] = nullptr; return false; }
      }
    }
    while (q.n_open >= GC_SEAM_BATCH) { lk.unlock(); std::this_thread::yield(); lk.lock(); }   // the open batch is full: the leader takes it in a moment
    struct InsideGuard { std::atomic<int>& c; InsideGuard(std::atomic<int>& x) : c(x) { c.fetch_add(1); } ~InsideGuard() { c.fetch_sub(1); } } inside_guard(q.inside);
    Seam::Waiter me; me.out = pHit; me.found = 0; me.done.store(0);
    const int buf = q.open_buf, idx = q.n_open++;
    gcore_dev::SeamIn& r = q.h_req[buf]->in[idx];
    memcpy(r.o, a_rayPos, 12); memcpy(r.d, a_rayDir, 12); r.t_near = t_near; r.t_far = t_far;
    q.h_req[buf]->out[idx].found = 0;
    q.waiter[buf][idx] = &me;
    if (!q.leader_active) {
      q.leader_active = true;
      while (q.n_open > 0) {                                            // lead: one launch per batch until nothing is pending
        if (q.inside.load() > q.n_open || q.last_n > q.n_open) {        // companions on their way in: a moment to file (a launch costs ~25 us)
          const int want = q.last_n;
          lk.unlock();
          const auto t0 = std::chrono::steady_clock::now();
          while (*(volatile int*)&q.n_open < want || q.inside.load() > *(volatile int*)&q.n_open) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(3)) break;
          }
          lk.lock();
        }
        const int b = q.open_buf, n = q.n_open;
        q.last_n = n; q.open_buf ^= 1; q.n_open = 0;                    // later callers file into the other buffer
        lk.unlock();
        const int rc = seam_launch(b, n);
        lk.lock();
        q.launches++; q.rays += (unsigned long long)n;
        for (int i = 0; i < n; i++) {                                   // hand the answers out (records are written only on a hit, :222-235)
          Seam::Waiter* w = q.waiter[b][i];
          w->found = (rc == 0) ? q.h_req[b]->out[i].found : 0;
          if (w->found) *w->out = q.h_req[b]->out[i].hit;
          w->done.store(1, std::memory_order_release);                  // (the waiter may leave its frame right after this store)
        }
      }
      q.leader_active = false;
      lk.unlock();
    } else {
      lk.unlock();
      int spins = 0;
      while (!me.done.load(std::memory_order_acquire)) {
        if (++spins > 2000) std::this_thread::yield();                  // more callers than cores: let the leader (or a filer) have this one
      }
    }
    return me.found != 0;
  } catch (...) { fail(GCORE_E_STATE, "gcore_closest_hit: internal error"); return false; }
}

int gcore_seam_stats(uint64_t* out2) {
  if (!out2) return fail(GCORE_E_ARG, "gcore_seam_stats: null");
  Seam& q = seam();
  std::lock_guard<SpinLock> lk(q.m);
  out2[0] = q.launches; out2[1] = q.rays;
  return 0;
}

int gcore_seam_server_stats(uint64_t* out2) {
  if (!out2) return fail(GCORE_E_ARG, "gcore_seam_server_stats: null");
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  out2[0] = seam().srv_starts; out2[1] = seam().srv_batches;
  return 0;
}

// ---------------------------------------------------------------------------------------------- analytic scene
int gcore_set_analytic_scene(const gcore_sphere* spheres, int nspheres, const gcore_cornell* box, const gcore_flatlight* fl) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (nspheres < 0 || nspheres > GC_MAX_SPHERES || (nspheres > 0 && !spheres)) return fail(GCORE_E_LIMIT, "gcore_set_analytic_scene: 0..96 spheres");
  // material ids index g_scn.mats[] unchecked in the kernels (the reference would raise Constraint_Error): refuse bad ones here
  for (int i = 0; i < nspheres; i++)
    if (spheres[i].mat < -1 || spheres[i].mat >= GC_MAX_MATERIALS) return fail(GCORE_E_ARG, "gcore_set_analytic_scene: sphere material id outside -1..31");
  if (box) for (int k = 0; k < 6; k++)
    if (box->mat_indices[k] < 0 || box->mat_indices[k] >= GC_MAX_MATERIALS) return fail(GCORE_E_ARG, "gcore_set_analytic_scene: box material id outside 0..31");
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
  mark_scene_dirty();
  return 0;
}

int gcore_set_materials(const gcore_material* materials, int n) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (n < 0 || n > GC_MAX_MATERIALS - 1 || (n > 0 && !materials)) return fail(GCORE_E_LIMIT, "gcore_set_materials: 0..31 materials");
  for (int i = 0; i < n; i++)
    if (materials[i].type < GCORE_MAT_NULL || materials[i].type > GCORE_MAT_PHONG) return fail(GCORE_E_ARG, "gcore_set_materials: unknown material type");
  memset(s.hs.mats, 0, sizeof(s.hs.mats));
  for (int i = 0; i < n; i++) {
    art::DMaterial& d = s.hs.mats[i];
    d.type = materials[i].type; memcpy(d.color, materials[i].color, 12); memcpy(d.color2, materials[i].color2, 12);
    d.param = materials[i].param; d.light = materials[i].light;
  }
  s.hs.nmaterials = n;
  mark_scene_dirty();
  return 0;
}

int gcore_set_light(const gcore_light* l) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!l) return fail(GCORE_E_ARG, "gcore_set_light: null");
  if (l->shape != GCORE_LIGHT_RECT && l->shape != GCORE_LIGHT_SPHERE) return fail(GCORE_E_ARG, "gcore_set_light: unknown shape");
  s.hs.light_shape = l->shape;
  memcpy(s.hs.l_boxMin, l->boxMin, 12); memcpy(s.hs.l_boxMax, l->boxMax, 12); memcpy(s.hs.l_normal, l->normal, 12);
  memcpy(s.hs.l_center, l->center, 12); s.hs.l_radius = l->radius; memcpy(s.hs.l_intensity, l->intensity, 12);
  s.hs.l_area = l->surfaceArea;
  mark_scene_dirty();
  return 0;
}

int gcore_set_camera(const gcore_camera* c) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!c) return fail(GCORE_E_ARG, "gcore_set_camera: null");
  memcpy(s.hs.cam_pos, c->pos, 12);
  memcpy(s.hs.cam_m, c->matrix, 48);                                   // rows 0..2 of the row-major 4x4
  mark_scene_dirty();
  return 0;
}

int gcore_set_mesh(const float* pos3, const float* nrm3, int nverts, const int32_t* tris3, int ntris, const float* bbox6, int mesh_mode) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (nverts < 0 || ntris < 0 || (ntris > 0 && (!pos3 || !nrm3 || !tris3))) return fail(GCORE_E_ARG, "gcore_set_mesh: bad arguments");
  if (mesh_mode != GCORE_MESH_REFERENCE_BF && mesh_mode != GCORE_MESH_NEAREST) return fail(GCORE_E_ARG, "gcore_set_mesh: bad mesh_mode");
  s.mesh_pos.assign(pos3, pos3 + 3 * (size_t)nverts);
  s.mesh_nrm.assign(nrm3, nrm3 + 3 * (size_t)nverts);
  s.mesh_tris.assign(tris3, tris3 + 3 * (size_t)ntris);
  if (bbox6) memcpy(s.mesh_bbox, bbox6, 24);
  else { memset(s.mesh_bbox, 0, 24); fold_bbox(s.mesh_pos, s.mesh_bbox); }
  return install_mesh(mesh_mode);
  GC_CATCH_INT
}

int gcore_set_mesh_vsgf(const char* path, const float* matrix16, int mesh_mode) {      // geometry.adb:499-609
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (!path || !matrix16) return fail(GCORE_E_ARG, "gcore_set_mesh_vsgf: null argument");
  if (mesh_mode != GCORE_MESH_REFERENCE_BF && mesh_mode != GCORE_MESH_NEAREST) return fail(GCORE_E_ARG, "gcore_set_mesh_vsgf: bad mesh_mode");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(GCORE_E_IO, std::string("cannot open ") + path);
  struct { int64_t fileSizeInBytes; int32_t verticesNum, indicesNum, materialsNum, flags; } hd;   // :499-507
  if (fread(&hd, 24, 1, f) != 1 || hd.verticesNum <= 0 || hd.indicesNum < 0) { fclose(f); return fail(GCORE_E_IO, "bad .vsgf header"); }
  // the counts come from an untrusted file: they must fit the bytes that are really there before anything is allocated
  fseek(f, 0, SEEK_END);
  const long long real_size = ftell(f);
  fseek(f, 24, SEEK_SET);
  const int vn = hd.verticesNum, tn = hd.indicesNum / 3;
  const long long need = 24 + (long long)vn * (16 + 16 + 8 + (hd.flags != 0 ? 16 : 0)) + (long long)tn * (12 + 4);
  if (real_size < need) { fclose(f); return fail(GCORE_E_IO, "truncated .vsgf file (header counts exceed the file size)"); }
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
  GC_CATCH_INT
}

int gcore_scene_init_cornell(const char* vsgf_path) {                                     // scene.adb:89-217
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
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
  s.use_committed = false;                                                               // the default build traces its own mesh
  M4 mrot = rotation(-M_PI_F / 6.0f, 0.0f, 1.0f, 0.0f);                                  // :194-206
  M4 mscale = ident(), mtans = ident();
  mtans.m[0][3] = -0.75f; mtans.m[1][3] = 0.1f; mtans.m[2][3] = 3.1f; mtans.m[3][3] = 1.0f;
  mscale.m[0][0] = 2.0f; mscale.m[1][1] = 2.0f; mscale.m[2][2] = 2.0f;
  M4 M = mul(mul(mtans, mrot), mscale);
  return gcore_set_mesh_vsgf(vsgf_path, &M.m[0][0], GCORE_MESH_REFERENCE_BF);
  GC_CATCH_INT
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

int gcore_get_committed_scene(float* pos3, int32_t* tris3, float* normals9, int32_t* mats, int* nverts, int* ntris) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  const FlatScene& f = s.flat;
  if (pos3 && !f.verts.empty()) memcpy(pos3, f.verts.data(), f.verts.size() * 4);
  if (tris3 && !f.tris.empty()) memcpy(tris3, f.tris.data(), f.tris.size() * 4);
  if (normals9 && !f.n9.empty()) memcpy(normals9, f.n9.data(), f.n9.size() * 4);
  if (mats && !f.mat.empty()) memcpy(mats, f.mat.data(), f.mat.size() * 4);
  if (nverts) *nverts = (int)f.verts.size() / 3;
  if (ntris) *ntris = (int)f.tris.size() / 3;
  return 0;
}

// ---------------------------------------------------------------------------------------------- batched scene queries
int gcore_closest_hit_batch(int n, const gcore_ray* rays, gcore_hit* hits) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (n < 0 || (n > 0 && (!rays || !hits))) return fail(GCORE_E_ARG, "gcore_closest_hit_batch: bad arguments");
  if (n == 0) return 0;
  if (int rc = sync_scene(0)) return rc;
  Device& dv = s.dev[0];
  gcore_ray* d_r = nullptr; gcore_hit* d_h = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d_r, sizeof(gcore_ray) * (size_t)n)) != cudaSuccess || (e = cudaMalloc(&d_h, sizeof(gcore_hit) * (size_t)n)) != cudaSuccess) {
    cudaFree(d_r); cudaFree(d_h); return fail_cuda(e, "cudaMalloc batch");
  }
  cudaMemcpyAsync(d_r, rays, sizeof(gcore_ray) * (size_t)n, cudaMemcpyHostToDevice, dv.stream);
  e = gcore_dev::launch_closest_batch(active_mesh_kind(), n, d_r, d_h, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_hit) * (size_t)n, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
  cudaFree(d_r); cudaFree(d_h);
  if (e != cudaSuccess) return fail_cuda(e, "closest batch");
  return 0;
  GC_CATCH_INT
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
  GC_TRY
  return render_units_locked(units, stats);
  GC_CATCH_INT
}

int gcore_get_counters(uint64_t* out4) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!out4) return fail(GCORE_E_ARG, "gcore_get_counters: null");
  unsigned long long c[6];
  if (int rc = read_counters(c)) return rc;
  out4[0] = c[0]; out4[1] = c[1]; out4[2] = c[2]; out4[3] = c[3];
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
  unsigned long long c[6];
  if (int rc = read_counters(c)) return rc;
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
  GC_TRY
  return read_frame_locked(hdr4, ldr, nullptr);
  GC_CATCH_INT
}

int gcore_render(const gcore_render_config* cfg, int width, int height, int units, float* hdr4, uint32_t* ldr, gcore_render_stats* stats) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  int rc;
  if ((rc = gcore_init_render(cfg))) return rc;
  if (width != s.width || height != s.height || !s.dev[0].d_accum) { if ((rc = gcore_resize_viewport(width, height))) return rc; }
  else if ((rc = alloc_accum())) return rc;                            // same viewport: clear like Resize_Viewport does (allocations are kept)
  gcore_render_stats st; memset(&st, 0, sizeof(st));
  if ((rc = render_units_locked(units, &st))) return rc;
  int launches = st.kernel_launches;
  if ((rc = read_frame_locked(hdr4, ldr, &launches))) return rc;
  Device& d0 = s.dev[0];
  CU(cudaSetDevice(d0.ordinal), "cudaSetDevice");
  CU(cudaEventRecord(d0.ev2, d0.stream), "event");
  CU(cudaEventSynchronize(d0.ev2), "event sync");
  float ms = 0.f; cudaEventElapsedTime(&ms, d0.ev0, d0.ev2);
  st.total_ms = ms; st.kernel_launches = launches;
  if (stats) *stats = st;
  return 0;
  GC_CATCH_INT
}

int gcore_load_frame(const float* hdr4, int spp) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (!hdr4 || spp < 0) return fail(GCORE_E_ARG, "gcore_load_frame: bad arguments");
  if (!s.dev[0].d_accum) return fail(GCORE_E_STATE, "gcore_load_frame: no viewport");
  const size_t px = (size_t)s.width * s.height;
  for (int d = 0; d < s.gpu_count; d++) {                              // every device picks its own tiles out of the full frame
    Device& dv = s.dev[d];
    if (int rc = ensure_device(d)) return rc;
    float4* tmp = nullptr;
    CU(cudaMalloc(&tmp, sizeof(float4) * px), "cudaMalloc frame");
    cudaError_t e = cudaMemcpyAsync(tmp, hdr4, sizeof(float4) * px, cudaMemcpyHostToDevice, dv.stream);
    RenderArgs a = make_args(d);
    if (e == cudaSuccess) e = gcore_dev::launch_load_accum(a, tmp, dv.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail_cuda(e, "load_frame");
  }
  s.spp = spp; s.palette_frame = false;
  return 0;
  GC_CATCH_INT
}

int gcore_primary_hits(int width, int height, gcore_hit* hits) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  GC_TRY
  if (width <= 0 || height <= 0 || !hits) return fail(GCORE_E_ARG, "gcore_primary_hits: bad arguments");
  if (int rc = sync_scene(0)) return rc;
  Device& dv = s.dev[0];
  const int sw = s.width, sh = s.height, srank = s.rank, sworld = s.world, sg = s.gpu_count;
  s.width = width; s.height = height; s.rank = 0; s.world = 1; s.gpu_count = 1;
  RenderArgs a = make_args(0);
  s.width = sw; s.height = sh; s.rank = srank; s.world = sworld; s.gpu_count = sg;
  a.nslots = a.ntiles * GC_TILE * GC_TILE; a.accum = nullptr;
  gcore_hit* d_h = nullptr;
  size_t px = (size_t)width * height;
  CU(cudaMalloc(&d_h, sizeof(gcore_hit) * px), "cudaMalloc hits");
  cudaError_t e = gcore_dev::launch_primary_hits(active_mesh_kind(), a, d_h, dv.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_h, sizeof(gcore_hit) * px, cudaMemcpyDeviceToHost, dv.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dv.stream);
  cudaFree(d_h);
  if (e != cudaSuccess) return fail_cuda(e, "primary hits");
  return 0;
  GC_CATCH_INT
}

int gcore_accum_device_ptr(void** d_accum, int64_t* nslots) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!s.dev[0].d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (d_accum) *d_accum = s.dev[0].d_accum;
  if (nslots) *nslots = s.dev[0].nslots;
  return 0;
}

int gcore_resolve_device(void* d_hdr4, void* d_ldr) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!s.dev[0].d_accum) return fail(GCORE_E_STATE, "no viewport");
  if (s.world != 1) return fail(GCORE_E_STATE, "gcore_resolve_device needs world == 1");
  return resolve_to((float4*)d_hdr4, (uint32_t*)d_ldr, nullptr);
}

int gcore_resolve_gathered_device(const void* d_all_accum, int spp, void* d_hdr4, void* d_ldr) {
  State& s = S();
  std::lock_guard<std::recursive_mutex> lk(s.mu);
  if (!d_all_accum || s.width <= 0) return fail(GCORE_E_ARG, "gcore_resolve_gathered_device: bad arguments");
  if (s.gpu_count != 1) return fail(GCORE_E_STATE, "gcore_resolve_gathered_device is the multi-process path (gpu_count == 1)");
  if (int rc = ensure_device(0)) return rc;
  RenderArgs a = make_args(0);
  CU(gcore_dev::launch_resolve_gathered(a, (const float4*)d_all_accum, spp > 0 ? spp : 1, (float4*)d_hdr4, (uint32_t*)d_ldr, s.dev[0].stream), "resolve_kernel");
  return 0;
}

}  // extern "C"
