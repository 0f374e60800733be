/* Plain-C consumer of include/tvc.h: proves the header is C (not C++), that libtvc.so can be bound with
 * nothing but dlopen, and exercises the entry points that need no GPU.  Built and run by
 * tests/test_abi_and_host.py with gcc. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "tvc.h"

#define LOAD(name)                                                   \
  *(void**)(&p_##name) = dlsym(h, #name);                            \
  if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

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
  LOAD(tvc_version) LOAD(tvc_status_string) LOAD(tvc_detector_params_default) LOAD(tvc_ctx_create)
  LOAD(tvc_ctx_destroy) LOAD(tvc_candidate_width) LOAD(tvc_query_row_bytes) LOAD(tvc_search)
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
  if (rc == TVC_OK) {            /* a B200 is present */
    if (!ctx || p_tvc_ctx_destroy(ctx) != TVC_OK) return 9;
    printf("abi ok (device)\n");
  } else {
    if (rc != TVC_ERR_NO_DEVICE || ctx != NULL) return 10;   /* no CPU fallback: the library says so */
    printf("abi ok (no device: %s)\n", p_tvc_status_string(rc));
  }
  dlclose(h);
  return 0;
}
