/* Plain-C consumer of include/tvc.h: proves the header is C (not C++), that libtvc.so can be bound with
 * nothing but dlopen, and exercises the entry points that need no GPU.  When a B200 is present it also
 * runs one small tvc_search (host buffers: 300 x 96 gallery with a planted duplicate, 7 queries, k = 5,
 * checked against a scalar C loop with the (similarity desc, index asc) rule) and tvc_k_occurrence over
 * the result (checked against a counting loop) - the hot path driven from C with no Python in the process.
 * Built and run by tests/test_abi_and_host.py (CPU leg) and tests/test_gpu_api.py (device leg) with gcc. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tvc.h"

#define LOAD(name)                                                   \
  *(void**)(&p_##name) = dlsym(h, #name);                            \
  if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

#define PN 300
#define PD 96
#define PM 7
#define PK 5

static float frand(unsigned* st) {
  *st = *st * 1664525u + 1013904223u;
  return (float)((*st >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

int main(int argc, char** argv) {
  if (argc < 2) return 64;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  int (*p_tvc_version)(void);
  const char* (*p_tvc_status_string)(int);
  void (*p_tvc_detector_params_default)(tvc_detector_params*);
  int (*p_tvc_ctx_create)(int, tvc_ctx**);
  int (*p_tvc_ctx_destroy)(tvc_ctx*);
  int (*p_tvc_candidate_width)(int32_t);
  int (*p_tvc_query_row_bytes)(int32_t);
  int (*p_tvc_search)(tvc_ctx*, tvc_gallery*, const void*, int, int64_t, int32_t, int32_t, float, uint32_t, float*,
                      int64_t*, void*);
  int (*p_tvc_gallery_create)(tvc_ctx*, const void*, int, int64_t, int32_t, int64_t, uint32_t, int64_t, void*,
                              tvc_gallery**);
  int (*p_tvc_gallery_destroy)(tvc_gallery*);
  int (*p_tvc_k_occurrence)(tvc_ctx*, const int64_t*, int64_t, int32_t, int64_t, int64_t, int32_t*, int, void*);
  const char* (*p_tvc_last_error)(tvc_ctx*);
  LOAD(tvc_version) LOAD(tvc_status_string) LOAD(tvc_detector_params_default) LOAD(tvc_ctx_create)
  LOAD(tvc_ctx_destroy) LOAD(tvc_candidate_width) LOAD(tvc_query_row_bytes) LOAD(tvc_search)
  LOAD(tvc_gallery_create) LOAD(tvc_gallery_destroy) LOAD(tvc_k_occurrence) LOAD(tvc_last_error)
  if (p_tvc_version() != TVC_VERSION) return 3;
  if (strcmp(p_tvc_status_string(TVC_OK), "ok") != 0) return 4;
  tvc_detector_params p;
  memset(&p, 0xff, sizeof p);
  p_tvc_detector_params_default(&p);
  if (p.n_variants != 5 || p.n_retrieval != 10 || p.n_generative != 3 || p.methods != 7u ||
      fabsf(p.detection_threshold - 0.5f) > 0 || fabsf(p.dedup_threshold - 0.95f) > 1e-6f ||
      fabsf(p.sigma_threshold - 0.30f) > 1e-6f)
    return 5;
  if (p_tvc_candidate_width(10) != 16 || p_tvc_candidate_width(26) != 32 || p_tvc_candidate_width(56) != 64 ||
      p_tvc_candidate_width(57) != 0)
    return 6;
  if (p_tvc_query_row_bytes(768) != 1536 || p_tvc_query_row_bytes(100) != 256) return 7;
  if (p_tvc_search(NULL, NULL, NULL, 0, 0, 0, 0, 0.f, 0u, NULL, NULL, NULL) != TVC_ERR_INVALID) return 8;
  tvc_ctx* ctx = NULL;
  const int rc = p_tvc_ctx_create(0, &ctx);
  if (rc == TVC_OK) {            /* a B200 is present: one small search + histogram, host buffers */
    if (!ctx) return 9;
    static float g[PN * PD], q[PM * PD], sim[PM * PK];
    static int64_t idx[PM * PK];
    static int32_t counts[PN], want_counts[PN];
    unsigned st = 12345u;
    for (int i = 0; i < PN; ++i) {
      double ss = 0.0;
      for (int j = 0; j < PD; ++j) { g[i * PD + j] = frand(&st); ss += (double)g[i * PD + j] * g[i * PD + j]; }
      for (int j = 0; j < PD; ++j) g[i * PD + j] = (float)(g[i * PD + j] / sqrt(ss));
    }
    memcpy(&g[200 * PD], &g[17 * PD], sizeof(float) * PD);        /* rows 17 and 200 tie exactly */
    for (int i = 0; i < PM; ++i)
      for (int j = 0; j < PD; ++j) q[i * PD + j] = g[(i * 17) * PD + j] + 0.05f * frand(&st);   /* query 1 sits on row 17 */
    tvc_gallery* gal = NULL;
    if (p_tvc_gallery_create(ctx, g, TVC_F32, PN, PD, 0, 0u, 0, NULL, &gal) != TVC_OK || !gal) {
      fprintf(stderr, "gallery_create: %s\n", p_tvc_last_error(ctx));
      return 11;
    }
    if (p_tvc_search(ctx, gal, q, TVC_F32, PM, PD, PK, -INFINITY, 0u, sim, idx, NULL) != TVC_OK) {
      fprintf(stderr, "search: %s\n", p_tvc_last_error(ctx));
      return 12;
    }
    memset(want_counts, 0, sizeof want_counts);
    for (int i = 0; i < PM; ++i) {
      float s[PN];
      for (int n = 0; n < PN; ++n) {
        float acc = 0.f;
        for (int j = 0; j < PD; ++j) acc += q[i * PD + j] * g[n * PD + j];
        s[n] = acc;
      }
      for (int r = 0; r < PK; ++r) {             /* selection by (similarity desc, index asc) */
        int best = -1;
        for (int n = 0; n < PN; ++n)
          if (!isnan(s[n]) && (best < 0 || s[n] > s[best])) best = n;
        const int64_t got = idx[i * PK + r];
        if (got < 0 || got >= PN) return 13;
        /* exact index unless the two similarities are within the 1e-3 band; similarities within 2e-3 */
        if (got != best && fabsf(s[got] - s[best]) > 1e-3f) return 14;
        if (fabsf(sim[i * PK + r] - s[best]) > 2e-3f) return 15;
        if (r > 0 && (sim[i * PK + r] > sim[i * PK + r - 1] ||
                      (sim[i * PK + r] == sim[i * PK + r - 1] && idx[i * PK + r] < idx[i * PK + r - 1])))
          return 16;                               /* ordered, ties by the lower index */
        want_counts[got] += 1;
        s[got] = NAN;
      }
    }
    if (idx[1 * PK + 0] != 17 || idx[1 * PK + 1] != 200 || sim[1 * PK] != sim[1 * PK + 1]) return 17;   /* the planted tie */
    if (p_tvc_k_occurrence(ctx, idx, PM, PK, 0, PN, counts, 1, NULL) != TVC_OK) return 18;
    if (memcmp(counts, want_counts, sizeof counts) != 0) return 19;
    if (p_tvc_search(ctx, gal, q, TVC_F32, PM, PD + 1, PK, -INFINITY, 0u, sim, idx, NULL) != TVC_ERR_INVALID) return 20;
    if (p_tvc_gallery_destroy(gal) != TVC_OK || p_tvc_ctx_destroy(ctx) != TVC_OK) return 21;
    printf("abi ok (device: search + k-occurrence from C)\n");
  } else {
    if (rc != TVC_ERR_NO_DEVICE || ctx != NULL) return 10;   /* no CPU fallback: the library says so */
    printf("abi ok (no device: %s)\n", p_tvc_status_string(rc));
  }
  dlclose(h);
  return 0;
}
