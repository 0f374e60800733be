// C ABI of libtvc.so (include/tvc.h): contexts, HBM-resident galleries, workspace management and
// the orchestration of the kernels.  Plain pointers and sizes; no exceptions cross the boundary.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <chrono>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "tvc_internal.h"

using namespace tvc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tvc_ctx {
  int device = 0;
  int sm_count = 148;
  // Locking (SURVEY §8b threading row: concurrent tvc_search on an immutable gallery): `mu` is held only for
  // short bookkeeping - the workspace map, the internal stream / event pools, the timing lists, the options,
  // a gallery's lazily built tensor maps.  A compute call holds the mutex of ITS stream's workspace for its
  // whole duration (enqueue order on a stream = lock order), never `mu`: calls on different streams overlap,
  // including the blocking synchronisation of the host-buffer path.
  std::mutex mu;
  EncodeTiledFn encode = nullptr;
  struct Ws {
    void* ptr = nullptr;
    size_t bytes = 0;
    std::mutex mu;
  };
  std::map<cudaStream_t, std::unique_ptr<Ws>> ws;
  std::vector<cudaStream_t> idle_streams;   // context-owned non-blocking streams for blocking host-buffer calls
  std::vector<cudaStream_t> own_streams;
  std::vector<cudaEvent_t> idle_events;
  int64_t debug_flags = 0;
  int64_t pair_min_rows = 4096;  // TVC_PAIR_MIN_ROWS overrides (0 = always, huge = never)
  // pacing of the CTA pairs of a wave (SearchPlan::pace): a producer may run pace_ahead blocks of pace_every
  // gallery tiles ahead of the slowest pair of its wave; pace_every = 0 switches it off (TVC_PACE_EVERY / _AHEAD)
  // resident-query kernel (tvc_gemm_topk_ts.cu): used when a unit spans at least this many 256-row gallery tiles
  // (the query tile is loaded into tensor memory once per unit).  OFF by default (INT64_MAX): bit-identical to the
  // pair kernel but at N = 64 - all that fits next to a 768-wide query tile in tensor memory - the MMA is bound by
  // the A read from TMEM (0.53x the pair kernel's rate, profiles/r2j_probe.log); kept as the measured experiment
  // and for its test (TVC_TS_MIN_TILES / option "ts_min_tiles" switch it on)
  int64_t ts_min_tiles = INT64_MAX;
  // pair kernel with a resident query tile: units of at least this many 256-row gallery tiles (the tile is
  // swapped once per unit, ~2 us); INT64_MAX = never (TVC_RQ_MIN_TILES / option "rq_min_tiles")
  int64_t rq_min_tiles = 64;
  int64_t rq_resident = 0;    // resident query k-blocks (5..7); 0 = by dimension (TVC_RQ_RESIDENT, measurements)
  int64_t pace_every = 8;     // measured on the bench workload (profiles/r2h_pace.log): 2..8 tiles x 1..4 blocks
  int64_t pace_ahead = 2;     // all give 98.7-99.2 ms per step against 102.0-102.6 unpaced; 16 x 3: 101.4-101.8
  // kernel (c): index streams of at least this many entries (over histograms beyond shared memory, up to 4 Mi bins)
  // take the bucketed two-pass path; 0 = whenever it applies, INT64_MAX = never (TVC_KOCC_PART_MIN / option)
  int64_t kocc_part_min = 1ll << 20;   // measured on a B200: 1 M entries 37 vs 45 us, 5 M 50 vs 84, 50 M 222 vs 456
  int64_t emb_trace_ptr = 0;     // debugging: device buffer for kernel (b) pipeline timestamps
  int64_t emb_generic = 0;       // 1: kernel (b) embedding mode always takes the one-warp-per-query kernel
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
  int64_t launches0 = 0;
};

struct tvc_gallery {
  tvc_ctx* ctx = nullptr;
  int64_t n = 0, cap = 0;
  int d = 0, d_pad = 0;
  int64_t offset = 0;
  uint32_t flags = 0;
  bool external = false;  // wraps caller-owned fp32 rows (tvc_gallery_wrap_f32): never freed, not searchable
  bool ipc = false;       // f32 was opened from a peer process (cudaIpcOpenMemHandle): closed on destroy
  std::vector<tvc_gallery*> parts;  // non-empty: a group of row shards (tvc_gallery_group_create)
  __nv_bfloat16* bf16 = nullptr;
  float* f32 = nullptr;
  CUtensorMap tmap;      // box [64 x 256]: single-CTA kernel
  CUtensorMap tmap128;   // box [64 x 128]: CTA-pair kernel (each CTA stages half a gallery tile)
  CUtensorMap tmap3;     // {64, rows, k-blocks} with box {64, 32, 4}: resident-query kernel
  bool tmap3_ok = false;
  int64_t tmap_rows = -1;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// tvc_last_error is per calling thread (like errno): concurrent calls cannot overwrite each other's message
thread_local std::string t_last_error;

int fail(tvc_ctx*, int status, const std::string& msg) {
  t_last_error = msg;
  return status;
}
int fail_cuda(tvc_ctx* ctx, cudaError_t e, const char* where) {
  cudaGetLastError();
  return fail(ctx, e == cudaErrorMemoryAllocation ? TVC_ERR_OOM : TVC_ERR_CUDA,
              std::string(where) + ": " + cudaGetErrorString(e));
}
#define TVC_CUDA(ctx, call)                                   \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call); \
  } while (0)

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

size_t elem_size(int dtype) { return dtype == TVC_F32 ? 4 : 2; }
size_t up256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// One compute call: the device guard, the stream the work is enqueued on and exclusive use of that stream's
// grow-only workspace.  Calls with device pointers run on the caller's stream (asynchronous, stream-ordered).
// A call with a host buffer blocks until its results are back, so it takes a context-owned non-blocking
// stream of its own, ordered after whatever is already queued on the caller's stream: threads calling with
// host buffers (the reference's ThreadPoolExecutor workers, src/pipeline.py:42,288,555-560) overlap their
// copies, kernels and waits instead of serialising on one workspace.
struct CallScope {
  tvc_ctx* ctx;
  DeviceGuard guard;
  cudaStream_t st;
  tvc_ctx::Ws* ws = nullptr;
  bool internal = false;
  int rc = TVC_OK;
  CallScope(tvc_ctx* c, void* stream, bool host_buffers)
      : ctx(c), guard(c->device), st(static_cast<cudaStream_t>(stream)) {
    if (!guard.ok) {
      rc = fail(ctx, TVC_ERR_CUDA, "cudaSetDevice failed");
      return;
    }
    cudaEvent_t ev = nullptr;
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      if (host_buffers) {
        cudaStream_t own = nullptr;
        if (!ctx->idle_streams.empty()) {
          own = ctx->idle_streams.back();
          ctx->idle_streams.pop_back();
        } else if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) == cudaSuccess) {
          ctx->own_streams.push_back(own);
        } else {
          cudaGetLastError();
          own = nullptr;
        }
        if (own) {
          if (!ctx->idle_events.empty()) {
            ev = ctx->idle_events.back();
            ctx->idle_events.pop_back();
          } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            ev = nullptr;
          }
          if (ev && cudaEventRecord(ev, st) == cudaSuccess && cudaStreamWaitEvent(own, ev, 0) == cudaSuccess) {
            st = own;
            internal = true;
          } else {           // could not order behind the caller's stream: stay on it
            cudaGetLastError();
            ctx->idle_streams.push_back(own);
          }
          if (ev) ctx->idle_events.push_back(ev);
        }
      }
      std::unique_ptr<tvc_ctx::Ws>& slot = ctx->ws[st];
      if (!slot) slot.reset(new (std::nothrow) tvc_ctx::Ws());
      ws = slot.get();
    }
    if (!ws) {
      rc = fail(ctx, TVC_ERR_OOM, "workspace record");
      return;
    }
    ws->mu.lock();
  }
  ~CallScope() {
    if (ws) ws->mu.unlock();
    if (internal) {
      std::lock_guard<std::mutex> lk(ctx->mu);
      ctx->idle_streams.push_back(st);
    }
  }
  CallScope(const CallScope&) = delete;
  CallScope& operator=(const CallScope&) = delete;
};

// Grow-only per-stream workspace; calls on one stream are stream-ordered so reuse is safe.
int get_ws(CallScope& cs, size_t bytes, uint8_t** out) {
  tvc_ctx* ctx = cs.ctx;
  cudaStream_t st = cs.st;
  tvc_ctx::Ws& w = *cs.ws;
  if (w.bytes < bytes) {
    static const bool trace = getenv("TVC_TRACE_WS") != nullptr;   // diagnostics: what a workspace growth costs
    const auto t0 = std::chrono::steady_clock::now();
    const size_t w_prev_bytes = w.bytes;
    if (w.ptr) {
      TVC_CUDA(ctx, cudaStreamSynchronize(st));
      TVC_CUDA(ctx, cudaFree(w.ptr));
      w.ptr = nullptr;
      w.bytes = 0;
    }
    const auto t1 = std::chrono::steady_clock::now();
    // Growing is expensive far beyond the allocation itself: cudaFree of the old block was measured at 0.7 s on a
    // B200 with a few GB mapped (it synchronises the device and unmaps), cudaMalloc at 30-70 ms (TVC_TRACE_WS=1
    // prints both).  So a workspace starts at 64 MB - every single-sample and few-thousand-row call of the drop-in
    // API fits without ever growing - and doubles from there.
    size_t want = bytes + (bytes >> 3);
    if (want < (64u << 20)) want = 64u << 20;
    if (want < 2 * w_prev_bytes) want = 2 * w_prev_bytes;
    TVC_CUDA(ctx, cudaMalloc(&w.ptr, want));
    w.bytes = want;
    if (trace) {
      const auto t2 = std::chrono::steady_clock::now();
      fprintf(stderr, "tvc: workspace of stream %p grows to %zu bytes: free %.3f ms, malloc %.3f ms\n", (void*)st, want,
              std::chrono::duration<double, std::milli>(t1 - t0).count(),
              std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
  }
  *out = static_cast<uint8_t*>(w.ptr);
  return TVC_OK;
}

int make_tmap(tvc_ctx* ctx, CUtensorMap* tm, const void* base, int64_t rows, int d_pad, int box_rows) {
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = ctx->encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                                 gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ctx, TVC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return TVC_OK;
}

int gallery_tmap(tvc_gallery* g) {
  std::lock_guard<std::mutex> lk(g->ctx->mu);   // concurrent first searches of one gallery
  if (g->tmap_rows == g->n) return TVC_OK;
  int rc = make_tmap(g->ctx, &g->tmap, g->bf16, g->n, g->d_pad, kBN);
  if (rc == TVC_OK) rc = make_tmap(g->ctx, &g->tmap128, g->bf16, g->n, g->d_pad, 128);
  if (rc == TVC_OK) {
    // the same rows seen as {k within a block, row, k-block}: one box = 32 rows x 4 k-blocks, landing as four
    // [32 x 64] swizzled slabs.  (The k-block stride is smaller than the row stride; if a driver refuses that,
    // searches stay on the pair kernel.)
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(kBK), static_cast<cuuint64_t>(g->n),
                                static_cast<cuuint64_t>(g->d_pad / kBK)};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(g->d_pad) * 2, static_cast<cuuint64_t>(kBK) * 2};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(kBK), 32, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    g->tmap3_ok = g->ctx->encode(&g->tmap3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, g->bf16, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    g->tmap_rows = g->n;
  }
  return rc;
}

int gallery_reserve(tvc_gallery* g, int64_t need, cudaStream_t st) {
  if (need <= g->cap) return TVC_OK;
  tvc_ctx* ctx = g->ctx;
  int64_t cap = g->cap + (g->cap >> 1);
  if (cap < need) cap = need;
  if (cap < 64) cap = 64;
  __nv_bfloat16* nb = nullptr;
  float* nf = nullptr;
  TVC_CUDA(ctx, cudaMalloc(&nb, static_cast<size_t>(cap) * g->d_pad * 2));
  if (!(g->flags & TVC_GALLERY_NO_MASTER)) {
    cudaError_t e = cudaMalloc(&nf, static_cast<size_t>(cap) * g->d * 4);
    if (e != cudaSuccess) {
      cudaFree(nb);
      return fail_cuda(ctx, e, "cudaMalloc(master)");
    }
  }
  cudaError_t ce = cudaSuccess;
  if (g->n > 0) {
    ce = cudaMemcpyAsync(nb, g->bf16, static_cast<size_t>(g->n) * g->d_pad * 2, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess && nf)
      ce = cudaMemcpyAsync(nf, g->f32, static_cast<size_t>(g->n) * g->d * 4, cudaMemcpyDeviceToDevice, st);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {   // the old allocation stays the gallery's; drop the new one
    cudaFree(nb);
    if (nf) cudaFree(nf);
    return fail_cuda(ctx, ce, "gallery_reserve: copy to the grown allocation");
  }
  if (g->bf16) cudaFree(g->bf16);
  if (g->f32) cudaFree(g->f32);
  g->bf16 = nb;
  g->f32 = nf;
  g->cap = cap;
  g->tmap_rows = -1;
  return TVC_OK;
}

// rows (host or device) -> device pointer usable by a kernel on `st` (staged in `stage` if host)
int to_device(tvc_ctx* ctx, const void* src, size_t bytes, uint8_t* stage, cudaStream_t st,
              const void** out) {
  if (is_device_ptr(src)) {
    *out = src;
    return TVC_OK;
  }
  TVC_CUDA(ctx, cudaMemcpyAsync(stage, src, bytes, cudaMemcpyHostToDevice, st));
  *out = stage;
  return TVC_OK;
}

cudaEvent_t get_event(tvc_ctx*) {
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

constexpr int64_t kMaxRowsPerLaunch = 1 << 20;
constexpr size_t kHostStageBytes = 64u << 20;   // staging window of tvc_gallery_append for host rows

// Stages a host array on the device (or passes a device pointer through); nullptr stays nullptr.
struct Stager {
  tvc_ctx* ctx;
  cudaStream_t st;
  uint8_t* ws;
  size_t off = 0;
  bool any_host = false;
  template <typename T>
  int in(const T* src, size_t count, const T** out) {
    *out = src;
    if (!src || count == 0 || is_device_ptr(src)) return TVC_OK;
    T* dst = reinterpret_cast<T*>(ws + off);
    off += up256(count * sizeof(T));
    any_host = true;
    TVC_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *out = dst;
    return TVC_OK;
  }
  template <typename T>
  T* out_buf(T* dst, size_t count, bool* staged) {
    *staged = false;
    if (!dst || count == 0 || is_device_ptr(dst)) return dst;
    T* b = reinterpret_cast<T*>(ws + off);
    off += up256(count * sizeof(T));
    any_host = true;
    *staged = true;
    return b;
  }
};


}  // namespace

extern "C" {

int tvc_version(void) { return TVC_VERSION; }

const char* tvc_status_string(int s) {
  switch (s) {
    case TVC_OK: return "ok";
    case TVC_ERR_INVALID: return "invalid argument";
    case TVC_ERR_CUDA: return "CUDA error";
    case TVC_ERR_NO_DEVICE: return "no sm_100 CUDA device (libtvc has no CPU fallback)";
    case TVC_ERR_UNSUPPORTED: return "unsupported request";
    case TVC_ERR_OOM: return "out of device memory";
    default: return "unknown status";
  }
}

void tvc_detector_params_default(tvc_detector_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->n_variants = 5;        /* DetectorConfig.num_text_variants, src/detector.py:183 */
  p->n_retrieval = 10;      /* DetectionConfig.retrieval_top_k, experiments/defenses/detector.py:29 */
  p->n_generative = 3;      /* DetectorConfig.num_reference_images, src/detector.py:188 */
  p->methods = 7u;
  p->aggregation = 0;
  p->w_text_variants = 0.4f;
  p->w_sd_reference = 0.4f;
  p->w_consistency = 0.2f;
  p->detection_threshold = 0.5f;
  p->voting = 1;
  for (int i = 0; i < 4; ++i) p->cc_weights[i] = 0.25f;
  p->cc_base_threshold = 0.5f;
  p->cc_adaptive = 1;
  p->dedup_threshold = 0.95f;
  p->sigma_threshold = 0.30f;
}

int tvc_ctx_create(int device, tvc_ctx** out) {
  if (!out) return TVC_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return TVC_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return TVC_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TVC_ERR_CUDA;
  if (prop.major != 10) return TVC_ERR_NO_DEVICE;  // tcgen05/TMEM kernels: sm_100 family only
  tvc_ctx* ctx = new (std::nothrow) tvc_ctx();
  if (!ctx) return TVC_ERR_OOM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  DeviceGuard guard(device);
  cudaFree(0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      fn == nullptr) {
    cudaGetLastError();
    delete ctx;
    return TVC_ERR_CUDA;
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  ctx->launches0 = launches_so_far();
  if (const char* e = getenv("TVC_PAIR_MIN_ROWS")) ctx->pair_min_rows = atoll(e);
  if (const char* e = getenv("TVC_TS_MIN_TILES")) ctx->ts_min_tiles = atoll(e);
  if (const char* e = getenv("TVC_RQ_MIN_TILES")) ctx->rq_min_tiles = atoll(e);
  if (const char* e = getenv("TVC_RQ_RESIDENT")) ctx->rq_resident = atoll(e);
  if (const char* e = getenv("TVC_KOCC_PART_MIN")) ctx->kocc_part_min = atoll(e);
  if (const char* e = getenv("TVC_PACE_EVERY")) ctx->pace_every = atoll(e);
  if (const char* e = getenv("TVC_PACE_AHEAD")) ctx->pace_ahead = atoll(e);
  *out = ctx;
  return TVC_OK;
}

int tvc_ctx_destroy(tvc_ctx* ctx) {
  if (!ctx) return TVC_OK;
  {
    DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->ws)
      if (kv.second && kv.second->ptr) cudaFree(kv.second->ptr);
    for (cudaStream_t s : ctx->own_streams) cudaStreamDestroy(s);
    for (cudaEvent_t e : ctx->idle_events) cudaEventDestroy(e);
    for (auto& pr : ctx->timed) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    for (auto& pr : ctx->event_pool) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  }
  delete ctx;
  return TVC_OK;
}

const char* tvc_last_error(tvc_ctx* ctx) { return ctx ? t_last_error.c_str() : "null context"; }

int64_t tvc_ctx_launch_count(tvc_ctx* ctx) { return ctx ? launches_so_far() - ctx->launches0 : 0; }

int tvc_ctx_set_option(tvc_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return TVC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (strcmp(name, "emb_trace_ptr") == 0) {
    ctx->emb_trace_ptr = value;
    return TVC_OK;
  }
  if (strcmp(name, "emb_generic") == 0) {
    ctx->emb_generic = value;
    return TVC_OK;
  }
  if (strcmp(name, "pair_min_rows") == 0) {
    ctx->pair_min_rows = value;
    return TVC_OK;
  }
  if (strcmp(name, "debug_flags") == 0) {
    ctx->debug_flags = value;
    return TVC_OK;
  }
  if (strcmp(name, "rq_resident") == 0) {
    ctx->rq_resident = (value >= 7 && value <= 9) ? value : 0;
    return TVC_OK;
  }
  if (strcmp(name, "rq_min_tiles") == 0) {
    ctx->rq_min_tiles = value;
    return TVC_OK;
  }
  if (strcmp(name, "ts_min_tiles") == 0) {
    ctx->ts_min_tiles = value;
    return TVC_OK;
  }
  if (strcmp(name, "kocc_part_min") == 0) {
    ctx->kocc_part_min = value < 0 ? 0 : value;
    return TVC_OK;
  }
  if (strcmp(name, "pace_every") == 0) {
    ctx->pace_every = value < 0 ? 0 : value;
    return TVC_OK;
  }
  if (strcmp(name, "pace_ahead") == 0) {
    ctx->pace_ahead = value < 1 ? 1 : value;
    return TVC_OK;
  }
  return fail(ctx, TVC_ERR_INVALID, std::string("unknown option ") + name);
}

int tvc_ctx_release_workspace(tvc_ctx* ctx, int64_t* freed_bytes) {
  if (!ctx) return TVC_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  std::vector<tvc_ctx::Ws*> all;      // (records stay: a running call holds a pointer to its own; only the memory goes)
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& kv : ctx->ws)
      if (kv.second) all.push_back(kv.second.get());
  }
  int64_t freed = 0;
  for (tvc_ctx::Ws* w : all) {        // lock order: a workspace mutex is never taken while holding ctx->mu
    std::lock_guard<std::mutex> wl(w->mu);
    if (!w->ptr) continue;
    TVC_CUDA(ctx, cudaDeviceSynchronize());   // nothing in flight may still read it
    TVC_CUDA(ctx, cudaFree(w->ptr));
    freed += static_cast<int64_t>(w->bytes);
    w->ptr = nullptr;
    w->bytes = 0;
  }
  if (freed_bytes) *freed_bytes = freed;
  return TVC_OK;
}

int tvc_ctx_set_timing(tvc_ctx* ctx, int enabled) {
  if (!ctx) return TVC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->timing = enabled != 0;
  return TVC_OK;
}

int tvc_ctx_last_search_kernel_ms(tvc_ctx* ctx, float* ms, int64_t* launches) {
  if (!ctx || !ms) return TVC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard guard(ctx->device);
  float total = 0.f;
  for (auto& pr : ctx->timed) {
    TVC_CUDA(ctx, cudaEventSynchronize(pr.second));
    float t = 0.f;
    TVC_CUDA(ctx, cudaEventElapsedTime(&t, pr.first, pr.second));
    total += t;
  }
  *ms = total;
  if (launches) *launches = static_cast<int64_t>(ctx->timed.size());
  for (auto& pr : ctx->timed) ctx->event_pool.push_back(pr);
  ctx->timed.clear();
  return TVC_OK;
}

// ------------------------------------------------------------------------------- gallery
int tvc_gallery_create(tvc_ctx* ctx, const void* rows, int dtype, int64_t n, int32_t d,
                       int64_t global_row_offset, uint32_t flags, int64_t capacity_hint, void* stream,
                       tvc_gallery** out) {
  if (!ctx || !out || d <= 0 || n < 0 || (n > 0 && !rows) || dtype < 0 || dtype > TVC_F16)
    return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_create: bad argument");
  if (n >= (1ll << 31)) return fail(ctx, TVC_ERR_UNSUPPORTED, "gallery shard larger than 2^31 rows");
  *out = nullptr;
  tvc_gallery* g = new (std::nothrow) tvc_gallery();
  if (!g) return TVC_ERR_OOM;
  g->ctx = ctx;
  g->d = d;
  g->d_pad = (d + kBK - 1) / kBK * kBK;
  g->offset = global_row_offset;
  g->flags = flags;
  int rc;
  {
    DeviceGuard guard(ctx->device);
    rc = gallery_reserve(g, capacity_hint > n ? capacity_hint : n, static_cast<cudaStream_t>(stream));
  }
  if (rc == TVC_OK && n > 0) rc = tvc_gallery_append(g, rows, dtype, n, stream);
  if (rc != TVC_OK) {
    tvc_gallery_destroy(g);
    return rc;
  }
  *out = g;
  return TVC_OK;
}

int tvc_gallery_append(tvc_gallery* g, const void* rows, int dtype, int64_t n, void* stream) {
  if (!g || n < 0 || (n > 0 && !rows) || dtype < 0 || dtype > TVC_F16) return TVC_ERR_INVALID;
  if (n == 0) return TVC_OK;
  tvc_ctx* ctx = g->ctx;
  if (g->external) return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_append: wrapped row view");
  if (g->n + n >= (1ll << 31)) return fail(ctx, TVC_ERR_UNSUPPORTED, "gallery shard larger than 2^31 rows");
  CallScope cs(ctx, stream, false);   // stays on the caller's stream: later searches on it see the rows
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  int rc = gallery_reserve(g, g->n + n, st);
  if (rc != TVC_OK) return rc;
  const size_t row_b = static_cast<size_t>(g->d) * elem_size(dtype);
  const bool host_src = !is_device_ptr(rows);
  const bool norm = (g->flags & TVC_GALLERY_NORMALIZE) != 0;
  if (!host_src) {
    TVC_CUDA(ctx, launch_prep_rows(rows, dtype, n, g->d, g->d_pad, norm, g->bf16 + static_cast<size_t>(g->n) * g->d_pad,
                                   g->f32 ? g->f32 + static_cast<size_t>(g->n) * g->d : nullptr, st));
  } else {
    // Host rows go through a BOUNDED staging window (<= kHostStageBytes of the grow-only workspace, not a
    // second copy of the whole gallery): copy and conversion of a window are stream-ordered, so the next
    // window's copy waits for the conversion that reads it and pinned sources stay fully asynchronous.
    int64_t win = static_cast<int64_t>(kHostStageBytes / row_b);
    if (win < 1) win = 1;
    if (win > n) win = n;
    uint8_t* ws;
    rc = get_ws(cs, up256(static_cast<size_t>(win) * row_b), &ws);
    if (rc != TVC_OK) return rc;
    for (int64_t r0 = 0; r0 < n; r0 += win) {
      const int64_t nr = n - r0 < win ? n - r0 : win;
      TVC_CUDA(ctx, cudaMemcpyAsync(ws, static_cast<const uint8_t*>(rows) + static_cast<size_t>(r0) * row_b,
                                    static_cast<size_t>(nr) * row_b, cudaMemcpyHostToDevice, st));
      TVC_CUDA(ctx, launch_prep_rows(ws, dtype, nr, g->d, g->d_pad, norm,
                                     g->bf16 + static_cast<size_t>(g->n + r0) * g->d_pad,
                                     g->f32 ? g->f32 + static_cast<size_t>(g->n + r0) * g->d : nullptr, st));
    }
  }
  g->n += n;
  g->tmap_rows = -1;
  if (host_src) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

int tvc_gallery_truncate(tvc_gallery* g, int64_t n) {
  if (!g || n < 0 || n > g->n) return TVC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(g->ctx->mu);   // (the tensor maps are rebuilt under the same mutex)
  g->n = n;
  g->tmap_rows = -1;
  return TVC_OK;
}

int tvc_gallery_move_row(tvc_gallery* g, int64_t src, int64_t dst, void* stream) {
  if (!g || g->external || src < 0 || dst < 0 || src >= g->n || dst >= g->n) return TVC_ERR_INVALID;
  if (src == dst) return TVC_OK;
  tvc_ctx* ctx = g->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, cudaMemcpyAsync(g->bf16 + dst * g->d_pad, g->bf16 + src * g->d_pad,
                                static_cast<size_t>(g->d_pad) * 2, cudaMemcpyDeviceToDevice, st));
  if (g->f32)
    TVC_CUDA(ctx, cudaMemcpyAsync(g->f32 + dst * g->d, g->f32 + src * g->d,
                                  static_cast<size_t>(g->d) * 4, cudaMemcpyDeviceToDevice, st));
  return TVC_OK;
}

int tvc_gallery_info(const tvc_gallery* g, int64_t* n, int32_t* d, int64_t* global_row_offset,
                     uint32_t* flags) {
  if (!g) return TVC_ERR_INVALID;
  if (n) *n = g->n;
  if (d) *d = g->d;
  if (global_row_offset) *global_row_offset = g->offset;
  if (flags) *flags = g->flags;
  return TVC_OK;
}

int tvc_gallery_device_ptrs(const tvc_gallery* g, const void** bf16_rows, int32_t* d_pad,
                            const float** f32_rows) {
  if (!g) return TVC_ERR_INVALID;
  if (bf16_rows) *bf16_rows = g->bf16;
  if (d_pad) *d_pad = g->d_pad;
  if (f32_rows) *f32_rows = g->f32;
  return TVC_OK;
}

int tvc_gallery_get_rows(tvc_gallery* g, const int64_t* idx, int64_t n, float* out, void* stream) {
  if (!g || n < 0 || (n > 0 && (!idx || !out))) return TVC_ERR_INVALID;
  if (n == 0) return TVC_OK;
  tvc_ctx* ctx = g->ctx;
  const bool idx_dev = is_device_ptr(idx), out_dev = is_device_ptr(out);
  CallScope cs(ctx, stream, !idx_dev || !out_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const size_t idx_b = up256(static_cast<size_t>(n) * 8), out_b = up256(static_cast<size_t>(n) * g->d * 4);
  uint8_t* ws = nullptr;
  if (!idx_dev || !out_dev) {
    int rc = get_ws(cs, idx_b + out_b, &ws);
    if (rc != TVC_OK) return rc;
  }
  const void* didx = idx;
  if (!idx_dev) {
    int rc = to_device(ctx, idx, static_cast<size_t>(n) * 8, ws, st, &didx);
    if (rc != TVC_OK) return rc;
  }
  float* dout = out_dev ? out : reinterpret_cast<float*>(ws + idx_b);
  TVC_CUDA(ctx, launch_gather_rows(g->f32, g->bf16, g->d, g->d_pad, static_cast<const int64_t*>(didx), n,
                                   g->n, dout, st));
  if (!out_dev) {
    TVC_CUDA(ctx, cudaMemcpyAsync(out, dout, static_cast<size_t>(n) * g->d * 4, cudaMemcpyDeviceToHost, st));
    TVC_CUDA(ctx, cudaStreamSynchronize(st));
  } else if (!idx_dev) {
    TVC_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TVC_OK;
}

int tvc_gallery_destroy(tvc_gallery* g) {
  if (!g) return TVC_OK;
  {
    DeviceGuard guard(g->ctx->device);
    if (g->ipc) {
      cudaDeviceSynchronize();
      if (g->f32) cudaIpcCloseMemHandle(g->f32);
    } else if (!g->external) {
      cudaDeviceSynchronize();
      if (g->bf16) cudaFree(g->bf16);
      if (g->f32) cudaFree(g->f32);
    }
  }
  delete g;
  return TVC_OK;
}

int tvc_gallery_wrap_f32(tvc_ctx* ctx, const float* device_rows, int64_t n, int32_t d,
                         int64_t global_row_offset, tvc_gallery** out) {
  if (!ctx || !out || d <= 0 || n < 0 || (n > 0 && !device_rows))
    return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_wrap_f32: bad argument");
  if (n > 0 && !is_device_ptr(device_rows))
    return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_wrap_f32: rows must be device memory");
  tvc_gallery* g = new (std::nothrow) tvc_gallery();
  if (!g) return TVC_ERR_OOM;
  g->ctx = ctx;
  g->n = g->cap = n;
  g->d = d;
  g->d_pad = (d + kBK - 1) / kBK * kBK;
  g->offset = global_row_offset;
  g->flags = 0;
  g->external = true;
  g->f32 = const_cast<float*>(device_rows);
  *out = g;
  return TVC_OK;
}

int tvc_gallery_export_ipc(tvc_gallery* g, void* handle) {
  if (!g || !handle || g->external || !g->f32 || !g->parts.empty())
    return fail(g ? g->ctx : nullptr, TVC_ERR_INVALID, "tvc_gallery_export_ipc: needs an owned gallery with an fp32 master");
  static_assert(sizeof(cudaIpcMemHandle_t) == TVC_IPC_HANDLE_BYTES, "handle size");
  DeviceGuard guard(g->ctx->device);
  TVC_CUDA(g->ctx, cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle), g->f32));
  return TVC_OK;
}

int tvc_gallery_import_ipc(tvc_ctx* ctx, const void* handle, int64_t n, int32_t d, int64_t global_row_offset,
                           tvc_gallery** out) {
  if (!ctx || !handle || !out || n < 0 || d <= 0) return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_import_ipc: bad argument");
  DeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  TVC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  tvc_gallery* g = new (std::nothrow) tvc_gallery();
  if (!g) return TVC_ERR_OOM;
  g->ctx = ctx;
  g->n = g->cap = n;
  g->d = d;
  g->d_pad = (d + kBK - 1) / kBK * kBK;
  g->offset = global_row_offset;
  g->external = true;
  g->ipc = true;
  g->f32 = static_cast<float*>(p);
  *out = g;
  return TVC_OK;
}

int tvc_gallery_group_create(tvc_ctx* ctx, tvc_gallery** parts, int32_t n_parts, tvc_gallery** out) {
  if (!ctx || !parts || !out || n_parts < 1 || n_parts > kMaxParts)
    return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_group_create: 1..16 parts");
  for (int i = 0; i < n_parts; ++i)
    if (!parts[i] || parts[i]->ctx != ctx || parts[i]->d != parts[0]->d || !parts[i]->parts.empty())
      return fail(ctx, TVC_ERR_INVALID, "tvc_gallery_group_create: parts must be plain galleries of one dimension");
  tvc_gallery* g = new (std::nothrow) tvc_gallery();
  if (!g) return TVC_ERR_OOM;
  g->ctx = ctx;
  g->d = parts[0]->d;
  g->d_pad = parts[0]->d_pad;
  g->external = true;  // owns nothing: the parts stay alive with their owners
  g->parts.assign(parts, parts + n_parts);
  *out = g;
  return TVC_OK;
}

static void fill_row_source(RowSource* rs, const tvc_gallery* g) {
  memset(rs, 0, sizeof(*rs));
  rs->d_pad = g->d_pad;
  if (g->parts.empty()) {
    rs->nparts = 1;
    rs->f32[0] = g->f32;
    rs->bf16[0] = g->bf16;
    rs->off[0] = g->offset;
    rs->n[0] = g->n;
    return;
  }
  rs->nparts = static_cast<int>(g->parts.size());
  for (int i = 0; i < rs->nparts; ++i) {
    rs->f32[i] = g->parts[i]->f32;
    rs->bf16[i] = g->parts[i]->bf16;
    rs->off[i] = g->parts[i]->offset;
    rs->n[i] = g->parts[i]->n;
  }
}

// ------------------------------------------------------------------------------- search
// candidates-only mode of search_chunk (sharded search, phase 1)
struct CandOut {
  const ScatterSpec* sc = nullptr;   // scatter to the slice owners, or
  float* val = nullptr;              // local [m_total, kp] lists (device)
  int64_t* idx = nullptr;
  int64_t m_total = 0;
};

static int search_chunk(CallScope& cs, tvc_gallery* g, const void* queries, bool q_dev, int q_dtype,
                        int64_t m, int64_t row0, int32_t k, float threshold, uint32_t flags,
                        float* out_sim, bool sim_dev, int64_t* out_idx, bool idx_dev,
                        const CandOut* cand = nullptr) {
  tvc_ctx* ctx = cs.ctx;
  cudaStream_t st = cs.st;
  const int d = g->d, d_pad = g->d_pad;
  // CTA pairs (cta_group::2) pay off once there are enough 256-row query tiles to feed 74 pairs
  bool pair = m >= ctx->pair_min_rows && (ctx->sm_count % 2) == 0;
  SearchPlan plan = make_search_plan(m, g->n, d_pad, k, ctx->sm_count, pair);
  plan.debug = static_cast<int>(ctx->debug_flags);
  plan.skip_self = (flags & TVC_SEARCH_SKIP_SELF) ? 1 : 0;
  plan.self_offset = row0 - g->offset;  // query row i <-> global gallery row i
  const bool prepared = (flags & TVC_SEARCH_PREPARED_Q) != 0;   // queries = bf16 [m, d_pad] made by tvc_prepare_queries
  if (prepared && !cand) return fail(ctx, TVC_ERR_UNSUPPORTED, "prepared queries serve tvc_search_candidates only");
  const size_t q_in_b = (q_dev || prepared) ? 0 : up256(static_cast<size_t>(m) * d * elem_size(q_dtype));
  const size_t q_bf_b = prepared ? 0 : up256(static_cast<size_t>(m) * d_pad * 2);
  const size_t q_f32_b = (g->f32 && !cand) ? up256(static_cast<size_t>(m) * d * 4) : 0;
  const int64_t full_rows = static_cast<int64_t>(plan.full_tiles) * (plan.pair ? 2 * kBM : kBM);
  const size_t cand_n = static_cast<size_t>(m) * plan.splits * plan.kp;   // (upper bound: unsplit rows own one list)
  const size_t cand_b = up256(cand_n * 4);
  const size_t os_b = (sim_dev || cand) ? 0 : up256(static_cast<size_t>(m) * k * 4);
  const size_t oi_b = (idx_dev || cand) ? 0 : up256(static_cast<size_t>(m) * k * 8);
  // pacing counters: only where the pairs of a wave stream more gallery than the L2 keeps
  size_t pace_b = 0;
  {
    const int64_t every = ctx->pace_every, ahead = ctx->pace_ahead;
    const size_t stream_b = static_cast<size_t>(plan.n_tiles) * kBN * d_pad * 2;
    if (plan.pair && plan.full_tiles > 0 && every > 0 && ahead > 0 && stream_b > (48u << 20) &&
        plan.n_tiles > every * (ahead + 1)) {
      plan.pace_every = static_cast<int>(every);
      plan.pace_ahead = static_cast<int>(ahead);
      plan.pace_blocks = static_cast<int>((plan.n_tiles + every - 1) / every);
      const int waves = plan.full_tiles / (plan.grid / 2);
      pace_b = up256(static_cast<size_t>(waves) * plan.pace_blocks * 4);
    }
  }
  uint8_t* ws;
  int rc = get_ws(cs, q_in_b + q_bf_b + q_f32_b + 2 * cand_b + os_b + oi_b + pace_b, &ws);
  if (rc != TVC_OK) return rc;
  uint8_t* p = ws;
  if (pace_b) {
    plan.pace = reinterpret_cast<unsigned int*>(p);
    p += pace_b;
    TVC_CUDA(ctx, cudaMemsetAsync(plan.pace, 0, pace_b, st));
  }
  uint8_t* q_in = p; p += q_in_b;
  __nv_bfloat16* q_bf = reinterpret_cast<__nv_bfloat16*>(p); p += q_bf_b;
  float* q_f32 = q_f32_b ? reinterpret_cast<float*>(p) : nullptr; p += q_f32_b;
  float* cand_val = reinterpret_cast<float*>(p); p += cand_b;
  int32_t* cand_idx = reinterpret_cast<int32_t*>(p); p += cand_b;
  float* d_sim = sim_dev ? out_sim : reinterpret_cast<float*>(p); p += os_b;
  int64_t* d_idx = idx_dev ? out_idx : reinterpret_cast<int64_t*>(p); p += oi_b;

  if (prepared) {
    q_bf = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(queries));
  } else {
    const void* q_src = queries;
    if (!q_dev) {
      rc = to_device(ctx, queries, static_cast<size_t>(m) * d * elem_size(q_dtype), q_in, st, &q_src);
      if (rc != TVC_OK) return rc;
    }
    TVC_CUDA(ctx, launch_prep_rows(q_src, q_dtype, m, d, d_pad, (flags & TVC_SEARCH_NORMALIZE_Q) != 0, q_bf,
                                   q_f32, st));
  }
  CUtensorMap tq;
  rc = make_tmap(ctx, &tq, q_bf, m, d_pad, kBM);
  if (rc != TVC_OK) return rc;
  rc = gallery_tmap(g);
  if (rc != TVC_OK) return rc;
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  bool timing;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    timing = ctx->timing;
    if (timing && !ctx->event_pool.empty()) {
      ev = ctx->event_pool.back();
      ctx->event_pool.pop_back();
    }
  }
  if (timing) {
    if (!ev.first) {
      ev.first = get_event(ctx);
      ev.second = get_event(ctx);
    }
    TVC_CUDA(ctx, cudaEventRecord(ev.first, st));
  }
  // shortest unit of the plan, in 256-row gallery tiles
  const int64_t unit_tiles = plan.rem_tiles > 0 ? plan.tiles_per_split : plan.n_tiles;
  const bool resident_q = plan.pair && g->tmap3_ok && plan.kblocks <= ts_max_kblocks() && unit_tiles >= ctx->ts_min_tiles;
  if (resident_q)
    TVC_CUDA(ctx, launch_gemm_topk_ts(g->tmap3, q_bf, plan, cand_val, cand_idx, st));
  else if (plan.pair && unit_tiles >= ctx->rq_min_tiles && plan.kblocks <= 32)
    TVC_CUDA(ctx, launch_gemm_topk_pair_rq(tq, g->tmap128, plan, cand_val, cand_idx, static_cast<int>(ctx->rq_resident), st));
  else if (plan.pair)
    TVC_CUDA(ctx, launch_gemm_topk_pair(tq, g->tmap128, plan, cand_val, cand_idx, st));
  else
    TVC_CUDA(ctx, launch_gemm_topk(tq, g->tmap, plan, cand_val, cand_idx, st));
  if (timing) {
    TVC_CUDA(ctx, cudaEventRecord(ev.second, st));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->timed.push_back(ev);
  }
  if (cand) {
    // sharded search: keep the kp best per row by GEMM score; re-ranking happens on the slice owner
    ScatterSpec sc{};
    if (cand->sc) {
      sc = *cand->sc;
      if (row0 != 0) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_search_candidates: scatter needs m <= 2^20 rows");
    }
    TVC_CUDA(ctx, launch_select_candidates(cand_val, cand_idx, m, full_rows, plan.splits, plan.kp, g->offset, sc,
                                           cand->val ? cand->val + row0 * plan.kp : nullptr,
                                           cand->idx ? cand->idx + row0 * plan.kp : nullptr, st));
    return TVC_OK;
  }
  TVC_CUDA(ctx, launch_rerank(cand_val, cand_idx, m, full_rows, plan.splits, plan.kp, k, q_f32, g->f32, d, threshold,
                              g->offset, d_sim, d_idx, st));
  if (!sim_dev)
    TVC_CUDA(ctx, cudaMemcpyAsync(out_sim, d_sim, static_cast<size_t>(m) * k * 4, cudaMemcpyDeviceToHost, st));
  if (!idx_dev)
    TVC_CUDA(ctx, cudaMemcpyAsync(out_idx, d_idx, static_cast<size_t>(m) * k * 8, cudaMemcpyDeviceToHost, st));
  return TVC_OK;
}

int tvc_search(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m, int32_t d,
               int32_t k, float threshold, uint32_t flags, float* out_sim, int64_t* out_idx,
               void* stream) {
  if (!ctx || !g || g->ctx != ctx) return fail(ctx, TVC_ERR_INVALID, "tvc_search: bad handle");
  if (m < 0 || (m > 0 && (!queries || !out_sim || !out_idx)) || q_dtype < 0 || q_dtype > TVC_F16)
    return fail(ctx, TVC_ERR_INVALID, "tvc_search: bad argument");
  if (g->external) return fail(ctx, TVC_ERR_INVALID, "tvc_search: row views / groups are not searchable");
  if (d != g->d) return fail(ctx, TVC_ERR_INVALID, "tvc_search: query dimension != gallery dimension");
  if (k < 1) return fail(ctx, TVC_ERR_INVALID, "tvc_search: k < 1");
  if (k > TVC_MAX_K) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_search: k > TVC_MAX_K");
  if (m == 0) return TVC_OK;
  const bool q_dev = is_device_ptr(queries), sim_dev = is_device_ptr(out_sim),
             idx_dev = is_device_ptr(out_idx);
  CallScope cs(ctx, stream, !q_dev || !sim_dev || !idx_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  if (g->n == 0) {
    // empty gallery: every slot unused
    std::vector<float> hs(static_cast<size_t>(m) * k, -INFINITY);
    std::vector<int64_t> hi(static_cast<size_t>(m) * k, -1);
    TVC_CUDA(ctx, cudaMemcpyAsync(out_sim, hs.data(), hs.size() * 4,
                                  sim_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, st));
    TVC_CUDA(ctx, cudaMemcpyAsync(out_idx, hi.data(), hi.size() * 8,
                                  idx_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, st));
    TVC_CUDA(ctx, cudaStreamSynchronize(st));
    return TVC_OK;
  }
  const size_t q_row_b = static_cast<size_t>(d) * elem_size(q_dtype);
  for (int64_t r0 = 0; r0 < m; r0 += kMaxRowsPerLaunch) {
    const int64_t mc = m - r0 < kMaxRowsPerLaunch ? m - r0 : kMaxRowsPerLaunch;
    const int rc = search_chunk(cs, g, static_cast<const uint8_t*>(queries) + r0 * q_row_b, q_dev, q_dtype,
                                mc, r0, k, threshold, flags, out_sim + r0 * k, sim_dev, out_idx + r0 * k,
                                idx_dev);
    if (rc != TVC_OK) return rc;
    // host staging in the shared workspace is reused by the next chunk
    if ((!q_dev || !sim_dev || !idx_dev) && r0 + mc < m) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if (!q_dev || !sim_dev || !idx_dev) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

int tvc_candidate_width(int32_t k) {
  if (k < 1 || k > TVC_MAX_K) return 0;
  return make_search_plan(256, 256, kBK, k, 148, false).kp;
}

int tvc_search_candidates(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m, int32_t d,
                          int32_t k, uint32_t flags, const tvc_scatter* scatter, float* cand_val,
                          int64_t* cand_idx, void* stream) {
  if (!ctx || !g || g->ctx != ctx) return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: bad handle");
  if (m < 0 || (m > 0 && !queries) || q_dtype < 0 || q_dtype > TVC_F16 || g->external || d != g->d)
    return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: bad argument");
  if (k < 1) return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: k < 1");
  if (k > TVC_MAX_K) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_search_candidates: k > TVC_MAX_K");
  if (!scatter && (!cand_val || !cand_idx) && m > 0)
    return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: neither scatter nor local lists given");
  if (scatter && (scatter->n_slices < 1 || scatter->n_slices > kMaxParts || scatter->rows_per_slice < 1 ||
                  scatter->slot < 0 || scatter->rows_per_slice * scatter->n_slices < m))
    return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: bad scatter description");
  if (!is_device_ptr(queries) || (!scatter && (!is_device_ptr(cand_val) || !is_device_ptr(cand_idx))))
    return fail(ctx, TVC_ERR_INVALID, "tvc_search_candidates: device pointers only");
  if (m == 0) return TVC_OK;
  CallScope cs(ctx, stream, false);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const int kp = tvc_candidate_width(k);
  ScatterSpec sc{};
  CandOut co;
  co.m_total = m;
  if (scatter) {
    sc.n_slices = scatter->n_slices;
    sc.slot = scatter->slot;
    sc.rows_per_slice = scatter->rows_per_slice;
    for (int i = 0; i < scatter->n_slices; ++i) {
      sc.val[i] = scatter->val[i];
      sc.idx[i] = reinterpret_cast<long long*>(scatter->idx[i]);
    }
    co.sc = &sc;
  } else {
    co.val = cand_val;
    co.idx = cand_idx;
  }
  if (g->n == 0) {
    // empty shard: every slot unused (written by the select kernel from an all-empty candidate list)
    uint8_t* ws;
    const size_t cb = up256(static_cast<size_t>(m) * kp * 4);
    int rc = get_ws(cs, 2 * cb, &ws);
    if (rc != TVC_OK) return rc;
    TVC_CUDA(ctx, cudaMemsetAsync(ws, 0xFF, 2 * cb, st));   // idx = -1 everywhere
    TVC_CUDA(ctx, launch_select_candidates(reinterpret_cast<float*>(ws), reinterpret_cast<int32_t*>(ws + cb), m, 0, 1,
                                           kp, g->offset, sc, co.val, co.idx, st));
    return TVC_OK;
  }
  const size_t q_row_b = (flags & TVC_SEARCH_PREPARED_Q) ? static_cast<size_t>(g->d_pad) * 2
                                                           : static_cast<size_t>(d) * elem_size(q_dtype);
  for (int64_t r0 = 0; r0 < m; r0 += kMaxRowsPerLaunch) {
    const int64_t mc = m - r0 < kMaxRowsPerLaunch ? m - r0 : kMaxRowsPerLaunch;
    const int rc = search_chunk(cs, g, static_cast<const uint8_t*>(queries) + r0 * q_row_b, true, q_dtype, mc, r0,
                                k, -INFINITY, flags, nullptr, true, nullptr, true, &co);
    if (rc != TVC_OK) return rc;
  }
  return TVC_OK;
}

int tvc_query_row_bytes(int32_t d) { return d > 0 ? (d + kBK - 1) / kBK * kBK * 2 : 0; }

int tvc_prepare_queries(tvc_ctx* ctx, const void* rows, int dtype, int64_t m, int32_t d, uint32_t flags,
                        int32_t n_dst, void* const* dst, int64_t dst_row0, void* stream) {
  if (!ctx || m < 0 || d <= 0 || dtype < 0 || dtype > TVC_F16 || n_dst < 1 || n_dst > kMaxParts || !dst ||
      dst_row0 < 0 || (m > 0 && !rows))
    return fail(ctx, TVC_ERR_INVALID, "tvc_prepare_queries: bad argument");
  if (m == 0) return TVC_OK;
  if (!is_device_ptr(rows)) return fail(ctx, TVC_ERR_INVALID, "tvc_prepare_queries: device pointers only");
  BcastSpec b{};
  b.n = n_dst;
  for (int i = 0; i < n_dst; ++i) {
    if (!dst[i]) return fail(ctx, TVC_ERR_INVALID, "tvc_prepare_queries: null destination");
    b.bf16[i] = static_cast<__nv_bfloat16*>(dst[i]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, launch_prep_rows_bcast(rows, dtype, m, d, (d + kBK - 1) / kBK * kBK,
                                       (flags & TVC_SEARCH_NORMALIZE_Q) != 0, b, dst_row0, nullptr, st));
  return TVC_OK;
}

int tvc_rerank_candidates(tvc_ctx* ctx, tvc_gallery* g, const float* queries, int64_t m, int32_t d,
                          int32_t parts, int32_t kp, const float* cand_val, const int64_t* cand_idx, int32_t k,
                          float threshold, float* out_sim, int64_t* out_idx, void* stream) {
  if (!ctx || !g || g->ctx != ctx) return fail(ctx, TVC_ERR_INVALID, "tvc_rerank_candidates: bad handle");
  if (m < 0 || d != g->d || parts < 1 || kp < 1 || kp > 64 || k < 1 || k > kp ||
      (m > 0 && (!queries || !cand_val || !cand_idx || !out_sim || !out_idx)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_rerank_candidates: bad argument");
  if (m == 0) return TVC_OK;
  if (!is_device_ptr(queries) || !is_device_ptr(cand_val) || !is_device_ptr(cand_idx) || !is_device_ptr(out_sim) ||
      !is_device_ptr(out_idx))
    return fail(ctx, TVC_ERR_INVALID, "tvc_rerank_candidates: device pointers only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(ctx->device);
  RowSource src;
  fill_row_source(&src, g);
  TVC_CUDA(ctx, launch_rerank_merged(cand_val, cand_idx, m, parts, kp, k, queries, src, d, threshold, out_sim,
                                     out_idx, st));
  return TVC_OK;
}

int tvc_exchange_merge(tvc_ctx* ctx, int64_t m, int32_t parts, int32_t kp, const float* cand_val,
                       const int64_t* cand_idx, int32_t n_dst, int64_t* const* req_dst, void* stream) {
  if (!ctx || m < 0 || parts < 1 || kp < 1 || kp > 64 || n_dst < 1 || n_dst > kMaxParts || !req_dst ||
      (m > 0 && (!cand_val || !cand_idx)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_exchange_merge: bad argument");
  if (m == 0) return TVC_OK;
  ReqDst dst{};
  dst.n = n_dst;
  for (int i = 0; i < n_dst; ++i) {
    if (!req_dst[i]) return fail(ctx, TVC_ERR_INVALID, "tvc_exchange_merge: null destination");
    dst.req[i] = reinterpret_cast<long long*>(req_dst[i]);
  }
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, launch_exchange_merge(cand_val, cand_idx, m, parts, kp, dst, static_cast<cudaStream_t>(stream)));
  return TVC_OK;
}

int tvc_exchange_rescore(tvc_ctx* ctx, tvc_gallery* shard, const float* q_f32, int32_t d, int32_t owners,
                         int64_t rows_per_slice, int64_t m_total, int32_t kp, const int64_t* req,
                         float* const* score_dst, void* stream) {
  if (!ctx || !shard || shard->ctx != ctx || !shard->parts.empty() || d != shard->d || owners < 1 ||
      owners > kMaxParts || rows_per_slice < 1 || m_total < 0 || m_total > rows_per_slice * owners || kp < 1 ||
      kp > 64 || !score_dst || (m_total > 0 && (!q_f32 || !req)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_exchange_rescore: bad argument");
  if (m_total == 0 || shard->n == 0) return TVC_OK;
  if (!shard->f32) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_exchange_rescore: shard without fp32 master");
  ScoreDst dst{};
  for (int i = 0; i < owners; ++i) {
    if (!score_dst[i]) return fail(ctx, TVC_ERR_INVALID, "tvc_exchange_rescore: null destination");
    dst.score[i] = score_dst[i];
  }
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, launch_exchange_rescore(req, q_f32, shard->f32, shard->offset, shard->n, d, owners, rows_per_slice,
                                        m_total, kp, dst, static_cast<cudaStream_t>(stream)));
  return TVC_OK;
}

int tvc_exchange_finalize(tvc_ctx* ctx, int64_t m, int32_t kp, int32_t k, float threshold, const int64_t* req,
                          const float* score, float* out_sim, int64_t* out_idx, void* stream) {
  if (!ctx || m < 0 || kp < 1 || kp > 64 || k < 1 || k > kp || (m > 0 && (!req || !score || !out_sim || !out_idx)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_exchange_finalize: bad argument");
  if (m == 0) return TVC_OK;
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, launch_exchange_finalize(req, score, m, kp, k, threshold, out_sim, out_idx,
                                         static_cast<cudaStream_t>(stream)));
  return TVC_OK;
}

int tvc_peer_copy(tvc_ctx* ctx, void* dst, const void* src, int64_t bytes, void* stream) {
  if (!ctx || bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(ctx, TVC_ERR_INVALID, "tvc_peer_copy: bad argument");
  if (bytes == 0) return TVC_OK;
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return TVC_OK;
}

// ------------------------------------------------------------------------------- retrieval metrics
int tvc_retrieval_metrics(tvc_ctx* ctx, const int64_t* topk_idx, int64_t q, int32_t k, const int64_t* rel_ptr,
                          const int64_t* rel_idx, int64_t n_rel, const int32_t* k_values, int32_t n_k, float* out,
                          void* stream) {
  if (!ctx || q < 0 || k < 1 || k > 64 || n_k < 0 || n_k > 8 || n_rel < 0 || (n_k > 0 && !k_values) ||
      (q > 0 && (!topk_idx || !rel_ptr || !out)) || (n_rel > 0 && !rel_idx))
    return fail(ctx, TVC_ERR_INVALID, "tvc_retrieval_metrics: bad argument");
  MetricKs ks{};
  ks.n = n_k;
  for (int i = 0; i < n_k; ++i) {
    if (k_values[i] < 1 || k_values[i] > k)
      return fail(ctx, TVC_ERR_INVALID, "tvc_retrieval_metrics: every K must lie in [1, k]");
    ks.k[i] = k_values[i];
  }
  if (q == 0) return TVC_OK;
  CallScope cs(ctx, stream, !is_device_ptr(topk_idx) || !is_device_ptr(rel_ptr) || (n_rel > 0 && !is_device_ptr(rel_idx)) ||
                               !is_device_ptr(out));
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const size_t Q = static_cast<size_t>(q), cols = 2 + 3 * static_cast<size_t>(n_k);
  uint8_t* ws;
  int rc = get_ws(cs, up256(Q * k * 8) + up256((Q + 1) * 8) + up256(static_cast<size_t>(n_rel) * 8 + 8) +
                               up256(Q * cols * 4), &ws);
  if (rc != TVC_OK) return rc;
  Stager sg{ctx, st, ws};
  const int64_t *d_topk, *d_ptr, *d_rel;
  if ((rc = sg.in(topk_idx, Q * k, &d_topk)) || (rc = sg.in(rel_ptr, Q + 1, &d_ptr)) ||
      (rc = sg.in(rel_idx, static_cast<size_t>(n_rel), &d_rel)))
    return rc;
  bool staged;
  float* d_out = sg.out_buf(out, Q * cols, &staged);
  TVC_CUDA(ctx, launch_retrieval_metrics(d_topk, q, k, d_ptr, d_rel, ks, d_out, st));
  if (staged) TVC_CUDA(ctx, cudaMemcpyAsync(out, d_out, Q * cols * 4, cudaMemcpyDeviceToHost, st));
  if (sg.any_host) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

// ------------------------------------------------------------------------------- peer buffers
int tvc_peer_alloc(tvc_ctx* ctx, int64_t bytes, void** ptr, void* handle) {
  if (!ctx || bytes <= 0 || !ptr || !handle) return fail(ctx, TVC_ERR_INVALID, "tvc_peer_alloc: bad argument");
  DeviceGuard guard(ctx->device);
  void* p = nullptr;
  if (cudaMalloc(&p, static_cast<size_t>(bytes)) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, TVC_ERR_OOM, "tvc_peer_alloc: out of device memory");
  }
  TVC_CUDA(ctx, cudaMemset(p, 0, static_cast<size_t>(bytes)));
  TVC_CUDA(ctx, cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle), p));
  *ptr = p;
  return TVC_OK;
}

int tvc_peer_open(tvc_ctx* ctx, const void* handle, void** ptr) {
  if (!ctx || !handle || !ptr) return fail(ctx, TVC_ERR_INVALID, "tvc_peer_open: bad argument");
  DeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  TVC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return TVC_OK;
}

int tvc_peer_close(tvc_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return TVC_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return TVC_OK;
}

int tvc_peer_free(tvc_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return TVC_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  TVC_CUDA(ctx, cudaFree(ptr));
  return TVC_OK;
}

int tvc_similarity_matrix(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m,
                          int32_t d, uint32_t flags, float* out, void* stream) {
  if (!ctx || !g || g->ctx != ctx) return fail(ctx, TVC_ERR_INVALID, "tvc_similarity_matrix: bad handle");
  if (m < 0 || (m > 0 && (!queries || !out)) || q_dtype < 0 || q_dtype > TVC_F16 || d != g->d || g->external)
    return fail(ctx, TVC_ERR_INVALID, "tvc_similarity_matrix: bad argument");
  if (m == 0 || g->n == 0) return TVC_OK;
  if (m >= (1ll << 31)) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_similarity_matrix: m too large");
  const bool q_dev = is_device_ptr(queries), out_dev = is_device_ptr(out);
  CallScope cs(ctx, stream, !q_dev || !out_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const size_t q_in_b = q_dev ? 0 : up256(static_cast<size_t>(m) * d * elem_size(q_dtype));
  const size_t q_bf_b = up256(static_cast<size_t>(m) * g->d_pad * 2);
  const size_t out_b = out_dev ? 0 : up256(static_cast<size_t>(m) * g->n * 4);
  uint8_t* ws;
  int rc = get_ws(cs, q_in_b + q_bf_b + out_b, &ws);
  if (rc != TVC_OK) return rc;
  const void* q_src = queries;
  if (!q_dev) {
    rc = to_device(ctx, queries, static_cast<size_t>(m) * d * elem_size(q_dtype), ws, st, &q_src);
    if (rc != TVC_OK) return rc;
  }
  __nv_bfloat16* q_bf = reinterpret_cast<__nv_bfloat16*>(ws + q_in_b);
  float* d_out = out_dev ? out : reinterpret_cast<float*>(ws + q_in_b + q_bf_b);
  TVC_CUDA(ctx, launch_prep_rows(q_src, q_dtype, m, d, g->d_pad, (flags & TVC_SEARCH_NORMALIZE_Q) != 0, q_bf,
                                 nullptr, st));
  CUtensorMap tq;
  rc = make_tmap(ctx, &tq, q_bf, m, g->d_pad, kBM);
  if (rc != TVC_OK) return rc;
  rc = gallery_tmap(g);
  if (rc != TVC_OK) return rc;
  TVC_CUDA(ctx, launch_gemm_store(tq, g->tmap, static_cast<int>(m), static_cast<int>(g->n), g->d_pad / kBK,
                                  d_out, g->n, ctx->sm_count, st));
  if (!out_dev) {
    TVC_CUDA(ctx, cudaMemcpyAsync(out, d_out, static_cast<size_t>(m) * g->n * 4, cudaMemcpyDeviceToHost, st));
    TVC_CUDA(ctx, cudaStreamSynchronize(st));
  } else if (!q_dev) {
    TVC_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TVC_OK;
}

int tvc_merge_topk(tvc_ctx* ctx, const float* in_sim, const int64_t* in_idx, int64_t m, int32_t parts,
                   int32_t k, float* out_sim, int64_t* out_idx, void* stream) {
  if (!ctx || m < 0 || parts < 1 || k < 1 || (m > 0 && (!in_sim || !in_idx || !out_sim || !out_idx)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_merge_topk: bad argument");
  if (m == 0) return TVC_OK;
  const bool dev = is_device_ptr(in_sim) && is_device_ptr(in_idx) && is_device_ptr(out_sim) &&
                   is_device_ptr(out_idx);
  CallScope cs(ctx, stream, !dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  if (dev) {
    TVC_CUDA(ctx, launch_merge_topk(in_sim, in_idx, m, parts, k, out_sim, out_idx, st));
    return TVC_OK;
  }
  if (is_device_ptr(in_sim) || is_device_ptr(in_idx) || is_device_ptr(out_sim) || is_device_ptr(out_idx))
    return fail(ctx, TVC_ERR_INVALID, "tvc_merge_topk: mix of host and device pointers");
  const size_t n_in = static_cast<size_t>(m) * parts * k, n_out = static_cast<size_t>(m) * k;
  const size_t b0 = up256(n_in * 4), b1 = up256(n_in * 8), b2 = up256(n_out * 4), b3 = up256(n_out * 8);
  uint8_t* ws;
  int rc = get_ws(cs, b0 + b1 + b2 + b3, &ws);
  if (rc != TVC_OK) return rc;
  float* ds = reinterpret_cast<float*>(ws);
  int64_t* di = reinterpret_cast<int64_t*>(ws + b0);
  float* os = reinterpret_cast<float*>(ws + b0 + b1);
  int64_t* oi = reinterpret_cast<int64_t*>(ws + b0 + b1 + b2);
  TVC_CUDA(ctx, cudaMemcpyAsync(ds, in_sim, n_in * 4, cudaMemcpyHostToDevice, st));
  TVC_CUDA(ctx, cudaMemcpyAsync(di, in_idx, n_in * 8, cudaMemcpyHostToDevice, st));
  TVC_CUDA(ctx, launch_merge_topk(ds, di, m, parts, k, os, oi, st));
  TVC_CUDA(ctx, cudaMemcpyAsync(out_sim, os, n_out * 4, cudaMemcpyDeviceToHost, st));
  TVC_CUDA(ctx, cudaMemcpyAsync(out_idx, oi, n_out * 8, cudaMemcpyDeviceToHost, st));
  TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

// ------------------------------------------------------------------------------- kernel (b)
static int check_params(tvc_ctx* ctx, const tvc_detector_params* p) {
  if (!p) return fail(ctx, TVC_ERR_INVALID, "null detector params");
  if (p->n_variants < 0 || p->n_variants > TVC_MAX_VARIANTS || p->n_retrieval < 0 ||
      p->n_retrieval > TVC_MAX_REFS || p->n_generative < 0 || p->n_generative > TVC_MAX_REFS)
    return fail(ctx, TVC_ERR_UNSUPPORTED, "V/R/G exceed TVC_MAX_VARIANTS/TVC_MAX_REFS");
  if (p->aggregation < 0 || p->aggregation > 3 || p->voting < 0 || p->voting > 2)
    return fail(ctx, TVC_ERR_INVALID, "bad aggregation/voting");
  return TVC_OK;
}

int tvc_consistency_sims(tvc_ctx* ctx, const tvc_detector_params* p, int64_t q, const float* s0,
                         const float* sv, const float* sr, const int32_t* r_cnt, const float* sg,
                         const int32_t* g_cnt, const float* sxv, float* scores, uint8_t* flags,
                         void* stream) {
  if (!ctx) return TVC_ERR_INVALID;
  int rc = check_params(ctx, p);
  if (rc != TVC_OK) return rc;
  if (q < 0 || (q > 0 && (!s0 || !scores || !flags)))
    return fail(ctx, TVC_ERR_INVALID, "tvc_consistency_sims: bad argument");
  if (q == 0) return TVC_OK;
  auto host = [](const void* ptr) { return ptr != nullptr && !is_device_ptr(ptr); };
  CallScope cs(ctx, stream, host(s0) || host(sv) || host(sr) || host(r_cnt) || host(sg) || host(g_cnt) || host(sxv) ||
                               host(scores) || host(flags));
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const size_t V = p->n_variants, R = p->n_retrieval, G = p->n_generative, X = V * (V - (V > 0)) / 2;
  const size_t Q = static_cast<size_t>(q);
  const size_t need = up256(Q * 4) + up256(Q * V * 4) + up256(Q * R * 4) + up256(Q * 4) + up256(Q * G * 4) +
                      up256(Q * 4) + up256(Q * X * 4) + up256(Q * TVC_NSCORES * 4) + up256(Q) + 4096;
  uint8_t* ws;
  rc = get_ws(cs, need, &ws);
  if (rc != TVC_OK) return rc;
  Stager sg_{ctx, st, ws};
  const float *d_s0, *d_sv, *d_sr, *d_sg, *d_sx;
  const int32_t *d_rc, *d_gc;
  if ((rc = sg_.in(s0, Q, &d_s0)) || (rc = sg_.in(sv, Q * V, &d_sv)) || (rc = sg_.in(sr, Q * R, &d_sr)) ||
      (rc = sg_.in(r_cnt, Q, &d_rc)) || (rc = sg_.in(sg, Q * G, &d_sg)) || (rc = sg_.in(g_cnt, Q, &d_gc)) ||
      (rc = sg_.in(sxv, Q * X, &d_sx)))
    return rc;
  bool st_scores, st_flags;
  float* d_scores = sg_.out_buf(scores, Q * TVC_NSCORES, &st_scores);
  uint8_t* d_flags = sg_.out_buf(flags, Q, &st_flags);
  TVC_CUDA(ctx, launch_consistency_sims(*p, q, d_s0, V ? d_sv : nullptr, R ? d_sr : nullptr, d_rc,
                                        G ? d_sg : nullptr, d_gc, X ? d_sx : nullptr, d_scores, d_flags, st));
  if (st_scores)
    TVC_CUDA(ctx, cudaMemcpyAsync(scores, d_scores, Q * TVC_NSCORES * 4, cudaMemcpyDeviceToHost, st));
  if (st_flags) TVC_CUDA(ctx, cudaMemcpyAsync(flags, d_flags, Q, cudaMemcpyDeviceToHost, st));
  if (sg_.any_host) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

int tvc_consistency_emb(tvc_ctx* ctx, const tvc_detector_params* p, int64_t q, int32_t d,
                        const float* img, const float* txt, const float* var,
                        tvc_gallery* ret_gallery, const int64_t* ret_idx, int32_t n_ret_cand,
                        const float* gen, const int32_t* g_cnt, tvc_gallery* gen_gallery,
                        const int64_t* gen_idx, int32_t n_gen_cand, float* scores, uint8_t* flags,
                        float* out_sv, float* out_sr, float* out_sg, void* stream) {
  if (!ctx) return TVC_ERR_INVALID;
  int rc = check_params(ctx, p);
  if (rc != TVC_OK) return rc;
  if (q < 0 || d <= 0 || (q > 0 && (!img || !txt || !scores || !flags)) || n_ret_cand < 0 || n_gen_cand < 0)
    return fail(ctx, TVC_ERR_INVALID, "tvc_consistency_emb: bad argument");
  if ((ret_idx && !ret_gallery) || (gen_idx && !gen_gallery))
    return fail(ctx, TVC_ERR_INVALID, "tvc_consistency_emb: index list without its gallery");
  if ((ret_gallery && ret_gallery->d != d) || (gen_gallery && gen_gallery->d != d))
    return fail(ctx, TVC_ERR_INVALID, "tvc_consistency_emb: gallery dimension mismatch");
  if (q == 0) return TVC_OK;
  const size_t V = p->n_variants, R = p->n_retrieval, G = p->n_generative, Q = static_cast<size_t>(q);
  const size_t D = static_cast<size_t>(d);
  const size_t need = 2 * up256(Q * D * 4) + up256(Q * V * D * 4) + up256(Q * n_ret_cand * 8) +
                      up256(Q * G * D * 4) + up256(Q * 4) + up256(Q * n_gen_cand * 8) +
                      up256(Q * TVC_NSCORES * 4) + up256(Q) + up256(Q * V * 4) + up256(Q * R * 4) +
                      up256(Q * G * 4) + 4096;
  // only pay for the staging we need: all-device callers get a minimal workspace
  const bool all_dev = is_device_ptr(img) && is_device_ptr(txt) && (!var || is_device_ptr(var)) &&
                       (!ret_idx || is_device_ptr(ret_idx)) && (!gen || is_device_ptr(gen)) &&
                       (!g_cnt || is_device_ptr(g_cnt)) && (!gen_idx || is_device_ptr(gen_idx)) &&
                       is_device_ptr(scores) && is_device_ptr(flags) &&
                       (!out_sv || is_device_ptr(out_sv)) && (!out_sr || is_device_ptr(out_sr)) &&
                       (!out_sg || is_device_ptr(out_sg));
  CallScope cs(ctx, stream, !all_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  // similarity lists handed from the gather/dot kernel to the statistics kernel
  const size_t X = V * (V > 0 ? V - 1 : 0) / 2;
  const size_t lists = 3 * up256(Q * 4) + up256(Q * V * 4) + up256(Q * R * 4) + up256(Q * G * 4) + up256(Q * X * 4);
  const size_t stage_bytes = all_dev ? 4096 : need;
  uint8_t* ws;
  rc = get_ws(cs, stage_bytes + lists, &ws);
  if (rc != TVC_OK) return rc;
  Stager sg_{ctx, st, ws};
  ConsistencyEmbArgs a{};
  {
    uint8_t* w = ws + stage_bytes;
    a.w_s0 = reinterpret_cast<float*>(w); w += up256(Q * 4);
    a.w_rcnt = reinterpret_cast<int32_t*>(w); w += up256(Q * 4);
    a.w_gcnt = reinterpret_cast<int32_t*>(w); w += up256(Q * 4);
    a.w_sv = reinterpret_cast<float*>(w); w += up256(Q * V * 4);
    a.w_sr = reinterpret_cast<float*>(w); w += up256(Q * R * 4);
    a.w_sg = reinterpret_cast<float*>(w); w += up256(Q * G * 4);
    a.w_sx = reinterpret_cast<float*>(w);
  }
  if ((rc = sg_.in(img, Q * D, &a.img)) || (rc = sg_.in(txt, Q * D, &a.txt)) ||
      (rc = sg_.in(var, Q * V * D, &a.var)) ||
      (rc = sg_.in(ret_idx, Q * static_cast<size_t>(n_ret_cand), &a.ret_idx)) ||
      (rc = sg_.in(gen, Q * G * D, &a.gen)) || (rc = sg_.in(g_cnt, Q, &a.g_cnt)) ||
      (rc = sg_.in(gen_idx, Q * static_cast<size_t>(n_gen_cand), &a.gen_idx)))
    return rc;
  if (V == 0) a.var = nullptr;
  if (G == 0) a.gen = nullptr;
  if (ret_gallery && ret_idx && R > 0 && n_ret_cand > 0) {
    fill_row_source(&a.ret, ret_gallery);
    a.n_ret_cand = n_ret_cand;
  } else {
    a.ret_idx = nullptr;
  }
  if (!a.gen && gen_gallery && gen_idx && G > 0 && n_gen_cand > 0) {
    fill_row_source(&a.genr, gen_gallery);
    a.n_gen_cand = n_gen_cand;
  } else {
    a.gen_idx = nullptr;
  }
  bool s0_, s1_, s2_, s3_, s4_;
  float* d_scores = sg_.out_buf(scores, Q * TVC_NSCORES, &s0_);
  uint8_t* d_flags = sg_.out_buf(flags, Q, &s1_);
  a.out_sv = sg_.out_buf(out_sv, Q * V, &s2_);
  a.out_sr = sg_.out_buf(out_sr, Q * R, &s3_);
  a.out_sg = sg_.out_buf(out_sg, Q * G, &s4_);
  a.trace = reinterpret_cast<unsigned long long*>(ctx->emb_trace_ptr);
  TVC_CUDA(ctx, launch_consistency_emb(*p, q, d, a, d_scores, d_flags, ctx->sm_count, ctx->emb_generic != 0, st));
  if (s0_) TVC_CUDA(ctx, cudaMemcpyAsync(scores, d_scores, Q * TVC_NSCORES * 4, cudaMemcpyDeviceToHost, st));
  if (s1_) TVC_CUDA(ctx, cudaMemcpyAsync(flags, d_flags, Q, cudaMemcpyDeviceToHost, st));
  if (s2_) TVC_CUDA(ctx, cudaMemcpyAsync(out_sv, a.out_sv, Q * V * 4, cudaMemcpyDeviceToHost, st));
  if (s3_) TVC_CUDA(ctx, cudaMemcpyAsync(out_sr, a.out_sr, Q * R * 4, cudaMemcpyDeviceToHost, st));
  if (s4_) TVC_CUDA(ctx, cudaMemcpyAsync(out_sg, a.out_sg, Q * G * 4, cudaMemcpyDeviceToHost, st));
  if (sg_.any_host) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

int tvc_reference_vector_rule(tvc_ctx* ctx, int64_t q, int32_t d, int32_t v, const float* img,
                              tvc_gallery* ret_gallery, const int64_t* ret_idx, int32_t k, const float* gen,
                              int32_t m, float sigma_threshold, float* out_s, float* out_ref, float* out_sigma,
                              uint8_t* flags, void* stream) {
  if (!ctx || q < 0 || d <= 0 || v < 1 || v > TVC_MAX_VARIANTS || k < 0 || m < 0 || k + m < 1 || k + m > 32 ||
      (q > 0 && (!img || !out_s || !out_sigma || !flags)) || (k > 0 && (!ret_gallery || !ret_idx)) || (m > 0 && !gen))
    return fail(ctx, TVC_ERR_INVALID, "tvc_reference_vector_rule: bad argument");
  if (ret_gallery && ret_gallery->d != d)
    return fail(ctx, TVC_ERR_INVALID, "tvc_reference_vector_rule: gallery dimension mismatch");
  if (q == 0) return TVC_OK;
  const size_t Q = static_cast<size_t>(q), D = static_cast<size_t>(d), V = static_cast<size_t>(v);
  const size_t need = up256(Q * D * 4) + up256(Q * V * k * 8) + up256(Q * V * m * D * 4) + up256(Q * V * 4) +
                      2 * up256(Q * 4) + up256(Q) + 4096;
  const bool all_dev = is_device_ptr(img) && (!ret_idx || is_device_ptr(ret_idx)) && (!gen || is_device_ptr(gen)) &&
                       is_device_ptr(out_s) && is_device_ptr(out_sigma) && is_device_ptr(flags) &&
                       (!out_ref || is_device_ptr(out_ref));
  CallScope cs(ctx, stream, !all_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  uint8_t* ws;
  const size_t valid_b = up256(Q * V);
  int rc = get_ws(cs, valid_b + (all_dev ? 4096 : need), &ws);
  if (rc != TVC_OK) return rc;
  uint8_t* valid_ws = ws;
  Stager sg{ctx, st, ws + valid_b};
  const float *d_img, *d_gen;
  const int64_t* d_idx;
  if ((rc = sg.in(img, Q * D, &d_img)) || (rc = sg.in(ret_idx, Q * V * k, &d_idx)) ||
      (rc = sg.in(gen, Q * V * m * D, &d_gen)))
    return rc;
  RowSource src{};
  if (k > 0) {
    fill_row_source(&src, ret_gallery);
    for (int i = 0; i < src.nparts; ++i)
      if (!src.f32[i]) return fail(ctx, TVC_ERR_UNSUPPORTED, "tvc_reference_vector_rule: gallery without fp32 master");
  }
  bool s0, s1, s2, s3;
  float* o_s = sg.out_buf(out_s, Q * V, &s0);
  float* o_ref = sg.out_buf(out_ref, Q, &s1);
  float* o_sig = sg.out_buf(out_sigma, Q, &s2);
  uint8_t* o_fl = sg.out_buf(flags, Q, &s3);
  TVC_CUDA(ctx, launch_reference_vector(q, d, v, d_img, src, d_idx, k, d_gen, m, sigma_threshold, o_s, o_ref, o_sig,
                                        o_fl, valid_ws, ctx->sm_count, st));
  if (s0) TVC_CUDA(ctx, cudaMemcpyAsync(out_s, o_s, Q * V * 4, cudaMemcpyDeviceToHost, st));
  if (s1) TVC_CUDA(ctx, cudaMemcpyAsync(out_ref, o_ref, Q * 4, cudaMemcpyDeviceToHost, st));
  if (s2) TVC_CUDA(ctx, cudaMemcpyAsync(out_sigma, o_sig, Q * 4, cudaMemcpyDeviceToHost, st));
  if (s3) TVC_CUDA(ctx, cudaMemcpyAsync(flags, o_fl, Q, cudaMemcpyDeviceToHost, st));
  if (sg.any_host) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

// ------------------------------------------------------------------------------- kernel (c)
int tvc_k_occurrence(tvc_ctx* ctx, const int64_t* idx, int64_t m, int32_t k, int64_t idx_base,
                     int64_t n_bins, int32_t* counts, int zero_first, void* stream) {
  if (!ctx || m < 0 || k < 1 || n_bins < 0 || (n_bins > 0 && !counts) || (m > 0 && !idx))
    return fail(ctx, TVC_ERR_INVALID, "tvc_k_occurrence: bad argument");
  if (n_bins == 0) return TVC_OK;
  const bool idx_dev = m == 0 || is_device_ptr(idx), cnt_dev = is_device_ptr(counts);
  CallScope cs(ctx, stream, !idx_dev || !cnt_dev);
  if (cs.rc != TVC_OK) return cs.rc;
  cudaStream_t st = cs.st;
  const size_t ib = up256(static_cast<size_t>(m) * k * 8), cb = up256(static_cast<size_t>(n_bins) * 4);
  // long streams over histograms beyond shared memory take the bucketed two-pass path (option "kocc_part_min");
  // a host stream is staged 256-byte aligned, so only a device stream can be misaligned for it
  int64_t part_min;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    part_min = ctx->kocc_part_min;
  }
  const size_t pb = k_occurrence_part_scratch_bytes(idx_dev ? idx : nullptr, m, k, n_bins, part_min);
  uint8_t* ws = nullptr;
  {
    int rc = get_ws(cs, 4096 + pb + ((!idx_dev || !cnt_dev) ? ib + cb : 0), &ws);
    if (rc != TVC_OK) return rc;
  }
  int* flag_scratch = reinterpret_cast<int*>(ws);
  ws += 4096;
  void* part_scratch = pb ? ws : nullptr;
  ws += pb;
  const int64_t* d_idx = idx;
  if (!idx_dev) {
    TVC_CUDA(ctx, cudaMemcpyAsync(ws, idx, static_cast<size_t>(m) * k * 8, cudaMemcpyHostToDevice, st));
    d_idx = reinterpret_cast<const int64_t*>(ws);
  }
  int32_t* d_cnt = counts;
  if (!cnt_dev) {
    d_cnt = reinterpret_cast<int32_t*>(ws + ib);
    if (!zero_first)
      TVC_CUDA(ctx, cudaMemcpyAsync(d_cnt, counts, static_cast<size_t>(n_bins) * 4, cudaMemcpyHostToDevice, st));
  }
  if (zero_first) TVC_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, static_cast<size_t>(n_bins) * 4, st));
  TVC_CUDA(ctx, launch_k_occurrence(d_idx, m, k, idx_base, n_bins, d_cnt, ctx->sm_count, flag_scratch, part_scratch, st));
  if (!cnt_dev)
    TVC_CUDA(ctx, cudaMemcpyAsync(counts, d_cnt, static_cast<size_t>(n_bins) * 4, cudaMemcpyDeviceToHost, st));
  if (!idx_dev || !cnt_dev) TVC_CUDA(ctx, cudaStreamSynchronize(st));
  return TVC_OK;
}

}  // extern "C"
