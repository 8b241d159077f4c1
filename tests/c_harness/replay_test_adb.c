/* replay_test_adb.c -- plain-C replay of the reference's main program (test.adb:20-79) against libgcore_b200.so.
 *
 * Test infrastructure: shows that the C ABI of include/gcore.h is usable with nothing but a C compiler and dlopen -- the
 * same calls, in the same order, that the Ada bodies of ada/ make through pragma Import (C, ...):
 *
 *   Scene.Init                      -> gcore_scene_init_cornell        (test.adb:32)
 *   Ray_Tracer.Init_Render(PT_MIS)  -> gcore_init_render               (test.adb:35)
 *   Ray_Tracer.Resize_Viewport      -> gcore_resize_viewport           (test.adb:36)
 *   loop: Ray_Tracer.Render_Pass    -> gcore_render_units(Threads_Num) + gcore_read_frame   (test.adb:50)
 *         Bitmap.SaveBMP            -> 24-bit BMP, bottom-up, B G R, no padding             (test.adb:63-69, bitmap.adb:31-85)
 *
 * usage: replay_test_adb <libgcore_b200.so> <pyramid2.vsgf> <out.bmp> [width height threads_num passes [gpu_count]]
 * gpu_count > 1: the same single host process drives that many devices (gcore_set_gpu_count) -- the BMP must not change.
 * prints one line: spp rays kernel_ms; exit code 0 on success.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gcore.h"

#define LOAD(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

static int save_bmp(const char* path, const uint32_t* data, int w, int h) {      /* bitmap.adb:31-85 */
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  unsigned char hd[54];
  memset(hd, 0, sizeof(hd));
  uint32_t size = 14 + 40 + (uint32_t)w * (uint32_t)h * 3u, off = 54, ihs = 40;
  hd[0] = 'B'; hd[1] = 'M';
  memcpy(hd + 2, &size, 4); memcpy(hd + 10, &off, 4); memcpy(hd + 14, &ihs, 4);
  int32_t ww = w, hh = h; memcpy(hd + 18, &ww, 4); memcpy(hd + 22, &hh, 4);
  hd[26] = 1; hd[28] = 24;
  fwrite(hd, 1, 54, f);
  for (int i = 0; i < w * h; i++) {
    unsigned char px[3] = {(unsigned char)((data[i] >> 16) & 255), (unsigned char)((data[i] >> 8) & 255), (unsigned char)(data[i] & 255)};
    fwrite(px, 1, 3, f);
  }
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s lib vsgf out.bmp [w h threads passes]\n", argv[0]); return 2; }
  int w = argc > 4 ? atoi(argv[4]) : 1024, hgt = argc > 5 ? atoi(argv[5]) : 768;
  int threads_num = argc > 6 ? atoi(argv[6]) : 28, passes = argc > 7 ? atoi(argv[7]) : 1, gpus = argc > 8 ? atoi(argv[8]) : 1;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(gcore_scene_init_cornell) LOAD(gcore_init_render) LOAD(gcore_resize_viewport) LOAD(gcore_render_units)
  LOAD(gcore_read_frame) LOAD(gcore_get_spp) LOAD(gcore_last_error) LOAD(gcore_set_gpu_count)
  char err[256];
  #define CK(x) do { int rc_ = (x); if (rc_ != 0) { p_gcore_last_error(err, 256); fprintf(stderr, "%s -> %d: %s\n", #x, rc_, err); return 1; } } while (0)
  if (gpus > 1) CK(p_gcore_set_gpu_count(gpus));
  CK(p_gcore_scene_init_cornell(argv[2]));
  gcore_render_config cfg = {GCORE_PT_MIS, 1, 8, 0x5EED0003u, {0.f, 0.f, 0.f}};      /* ray_tracer.ads:20-32 defaults */
  CK(p_gcore_init_render(&cfg));
  CK(p_gcore_resize_viewport(w, hgt));
  uint32_t* image = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)w * hgt);
  gcore_render_stats st;
  uint64_t rays = 0; float ms = 0.f;
  for (int p = 0; p < passes; p++) {                                                 /* test.adb:48-70 */
    CK(p_gcore_render_units(threads_num, &st));
    rays += st.rays; ms += st.kernel_ms;
    CK(p_gcore_read_frame(NULL, image));
    if (save_bmp(argv[3], image, w, hgt) != 0) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
  }
  printf("%d %llu %.3f\n", p_gcore_get_spp(), (unsigned long long)rays, ms);
  free(image);
  dlclose(h);
  return 0;
}
