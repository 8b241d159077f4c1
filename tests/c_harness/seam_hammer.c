/* seam_hammer.c -- T host threads call the legacy one-ray entry point gcore_closest_hit concurrently, the way the Ada tasks of
 * ray_tracer.adb:264-277 reach it through Scene.Find_Closest_Hit (scene_hydra_embree.adb:426-446).  Prints one JSON line with
 * the call rate; every thread also checks its answers against a single-threaded pass over the same rays.
 *   seam_hammer <libgcore_b200.so> <threads> <rays per thread> */
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int primIndex, geomIndex, instIndex; float t, normal[3], texCoord[2]; } HitCpp;
typedef void (*fn_void)(void);
typedef int (*fn_add)(const float*, int, const int*, int);
typedef bool (*fn_hit)(const float*, const float*, float, float, HitCpp*);
typedef int (*fn_stats)(unsigned long long*);

static fn_hit closest_hit;
static int per_thread;
static float* ray_o; static float* ray_d; static int* want_prim; static float* want_t;

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }
static float frand(unsigned* s, float lo, float hi) { return lo + (hi - lo) * (float)lcg(s) / 16777216.0f; }

static void* worker(void* arg) {
  long k = (long)arg, bad = 0;
  for (int i = k * per_thread; i < (k + 1) * per_thread; i++) {
    HitCpp h; memset(&h, 0, sizeof h); h.primIndex = -1;
    closest_hit(ray_o + 3 * i, ray_d + 3 * i, 0.0f, 100000.0f, &h);
    if (want_prim && (h.primIndex != want_prim[i] || (h.primIndex >= 0 && memcmp(&h.t, &want_t[i], 4) != 0))) bad++;
  }
  return (void*)bad;
}

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: seam_hammer lib threads rays_per_thread\n"); return 2; }
  void* L = dlopen(argv[1], RTLD_NOW);
  if (!L) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  int threads = atoi(argv[2]); per_thread = atoi(argv[3]);
  fn_void init = (fn_void)dlsym(L, "gcore_init_and_clear"), commit = (fn_void)dlsym(L, "gcore_commit_scene"), destroy = (fn_void)dlsym(L, "gcore_destroy");
  fn_add add = (fn_add)dlsym(L, "gcore_add_mesh_3f");
  closest_hit = (fn_hit)dlsym(L, "gcore_closest_hit");
  if (!init || !commit || !destroy || !add || !closest_hit) { fprintf(stderr, "missing symbol\n"); return 2; }
  /* 20 000 random triangles */
  const int ntris = 20000;
  unsigned seed = 12345u;
  float* verts = malloc(sizeof(float) * 9 * ntris); int* idx = malloc(sizeof(int) * 3 * ntris);
  for (int t = 0; t < ntris; t++) {
    float c[3] = {frand(&seed, -2, 2), frand(&seed, -2, 2), frand(&seed, -2, 2)};
    for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) verts[9 * t + 3 * v + a] = c[a] + frand(&seed, -0.15f, 0.15f);
    idx[3 * t] = 3 * t; idx[3 * t + 1] = 3 * t + 1; idx[3 * t + 2] = 3 * t + 2;
  }
  init(); add(verts, 3 * ntris, idx, 3 * ntris); commit();
  int n = threads * per_thread;
  ray_o = malloc(sizeof(float) * 3 * n); ray_d = malloc(sizeof(float) * 3 * n);
  for (int i = 0; i < n; i++) {
    float d[3], l = 0;
    for (int a = 0; a < 3; a++) { ray_o[3 * i + a] = frand(&seed, -3, 3); d[a] = frand(&seed, -1, 1); l += d[a] * d[a]; }
    l = 1.0f / sqrtf(l > 0 ? l : 1);
    for (int a = 0; a < 3; a++) ray_d[3 * i + a] = d[a] * l;
  }
  /* single-threaded pass: the expected answers (and the warm-up) */
  int* wp = malloc(sizeof(int) * n); float* wt = malloc(sizeof(float) * n);
  int nref = n < 4000 ? n : 4000;
  for (int i = 0; i < n; i++) { wp[i] = -2; wt[i] = 0; }
  for (int i = 0; i < nref; i++) { HitCpp h; memset(&h, 0, sizeof h); h.primIndex = -1; closest_hit(ray_o + 3 * i, ray_d + 3 * i, 0.0f, 100000.0f, &h); wp[i] = h.primIndex; wt[i] = h.t; }
  pthread_t* th = malloc(sizeof(pthread_t) * threads);
  fn_stats seam_stats = (fn_stats)dlsym(L, "gcore_seam_stats");
  unsigned long long s0[2] = {0, 0}, s1[2] = {0, 0};
  if (seam_stats) seam_stats(s0);
  double t0 = now();
  for (long k = 0; k < threads; k++) pthread_create(&th[k], NULL, worker, (void*)k);
  for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
  double dt = now() - t0;
  if (seam_stats) seam_stats(s1);
  /* correctness pass: same rays again from all threads, compared with the single-threaded answers where known */
  for (int i = nref; i < n; i++) { wp[i] = -2; }
  want_prim = NULL;
  long bad = 0, hits = 0;
  {
    /* compare the first nref rays threaded vs single-threaded */
    int* got = malloc(sizeof(int) * nref);
    (void)got;
    want_prim = wp; want_t = wt;
    int save = per_thread; per_thread = nref / threads;
    for (long k = 0; k < threads; k++) pthread_create(&th[k], NULL, worker, (void*)k);
    for (int k = 0; k < threads; k++) { void* r; pthread_join(th[k], &r); bad += (long)r; }
    per_thread = save;
    for (int i = 0; i < nref; i++) hits += wp[i] >= 0;
  }
  fn_stats server_stats = (fn_stats)dlsym(L, "gcore_seam_server_stats");
  unsigned long long sv[2] = {0, 0};
  if (server_stats) server_stats(sv);
  printf("{\"threads\": %d, \"rays\": %d, \"seconds\": %.4f, \"rays_per_s\": %.1f, \"us_per_ray_per_thread\": %.2f, \"mismatches\": %ld, \"hit_rate\": %.3f, \"batches\": %llu, \"rays_per_batch\": %.2f, \"resident_kernel_starts\": %llu, \"resident_kernel_batches\": %llu}\n",
         threads, n, dt, n / dt, 1e6 * dt * threads / n, bad, (double)hits / nref, s1[0] - s0[0], (double)(s1[1] - s0[1]) / (double)(s1[0] - s0[0] ? s1[0] - s0[0] : 1), sv[0], sv[1]);
  destroy();
  return bad ? 1 : 0;
}
