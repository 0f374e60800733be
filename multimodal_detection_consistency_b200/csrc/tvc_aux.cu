// Bandwidth-bound kernels around the GEMM: operand preparation, candidate re-rank / merge / exchange and
// kernel (c), the k-occurrence histogram.  128-bit loads, warp shuffles, shared-memory staging,
// warp-aggregated atomics (__match_any_sync: one RED per distinct bin a warp holds, used where the sampling
// pre-pass finds the stream repeating itself inside a warp) and, for index streams from 1 Mi entries, a bucketed
// two-pass path (copy-engine-fed bucket sort into 16-bit keys, then shared-memory atomics); no tensor cores.
#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "tvc_internal.h"
#include "tvc_ptx.cuh"

namespace tvc {

static std::atomic<int64_t> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// ------------------------------------------------------------------------------- prep_rows
// One warp per row: optional L2 normalisation in fp32, then bf16 [d_pad] (zero padded) and fp32 [d].
template <typename T>
__global__ void prep_rows_kernel(const T* __restrict__ rows, long long n, int d, int d_pad,
                                 int normalize, __nv_bfloat16* __restrict__ out_bf16,
                                 float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp0; r < n; r += nwarps) {
    const T* src = rows + r * d;
    float scale = 1.0f;
    if (normalize) {
      float ss = 0.f;
      for (int i = lane; i < d; i += 32) {
        const float x = to_f32(src[i]);
        ss = fmaf(x, x, ss);
      }
      ss = warp_sum(ss);
      scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    for (int i = lane; i < d_pad; i += 32) {
      const float x = i < d ? to_f32(src[i]) * scale : 0.f;
      if (out_bf16) out_bf16[r * d_pad + i] = __float2bfloat16_rn(x);
      if (out_f32 && i < d) out_f32[r * d + i] = x;
    }
  }
}

// Vectorised form for fp32 rows that need no normalisation (the query batches of the TVC flow):
// each lane converts 8 consecutive elements per trip - two 128-bit loads, one 128-bit bf16 store per
// destination, two 128-bit fp32 stores.  With several destinations the bf16 rows are BROADCAST: every
// pointer of `dst` is the query buffer of one rank of the box (own HBM or a peer's, mapped through
// CUDA IPC), so one rank converts its slice of the batch once and the stores over NVLink replace
// both the replicated conversion and an all-gather of the operand.
__global__ void __launch_bounds__(256)
prep_rows_f32v_kernel(const float* __restrict__ rows, long long n, int d, int d_pad, const BcastSpec dst,
                      long long dst_row0, float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int chunks = d_pad >> 3, live = d >> 3;   // d % 8 == 0 (launcher)
  for (long long r = warp0; r < n; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(rows + r * d);
    for (int c = lane; c < chunks; c += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (c < live) {
        a = src[2 * c];
        b = src[2 * c + 1];
        if (out_f32) {
          float4* o = reinterpret_cast<float4*>(out_f32 + r * d);
          o[2 * c] = a;
          o[2 * c + 1] = b;
        }
      }
      union {
        __nv_bfloat162 h[4];
        uint4 u;
      } pk;
      pk.h[0] = __floats2bfloat162_rn(a.x, a.y);
      pk.h[1] = __floats2bfloat162_rn(a.z, a.w);
      pk.h[2] = __floats2bfloat162_rn(b.x, b.y);
      pk.h[3] = __floats2bfloat162_rn(b.z, b.w);
      for (int t = 0; t < dst.n; ++t)
        reinterpret_cast<uint4*>(dst.bf16[t] + (dst_row0 + r) * d_pad)[c] = pk.u;
    }
  }
}

// ------------------------------------------------------------------------------- ordering
// Total order on (value, index): a is better when its value is larger, ties to the lower index.
__device__ __forceinline__ bool better(float v1, long long i1, float v2, long long i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

// Warp-wide: best (value, index) among the entries e < n that are strictly worse than the bound.
// `get(e, v, i)` fetches entry e; entries with i < 0 are empty.  Returns false when none is left.
template <typename Get>
__device__ __forceinline__ bool warp_next_best(int n, float bv, long long bi, Get get, float& ov,
                                               long long& oi) {
  const int lane = threadIdx.x & 31;
  float lv = -INFINITY;
  long long li = LLONG_MAX;
  for (int e = lane; e < n; e += 32) {
    float v;
    long long i;
    get(e, v, i);
    if (i >= 0 && better(bv, bi, v, i) && better(v, i, lv, li)) {
      lv = v;
      li = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(kFull, lv, o);
    const long long i2 = __shfl_xor_sync(kFull, li, o);
    if (better(v2, i2, lv, li)) {
      lv = v2;
      li = i2;
    }
  }
  ov = lv;
  oi = li;
  return li != LLONG_MAX;
}

__device__ __forceinline__ float warp_dot_f32(const float* __restrict__ a, const float* __restrict__ b,
                                              int d) {
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  if ((d & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int i = lane; i < (d >> 2); i += 32) {
      const float4 x = a4[i], y = b4[i];
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
      acc = fmaf(x.z, y.z, acc);
      acc = fmaf(x.w, y.w, acc);
    }
  } else {
    for (int i = lane; i < d; i += 32) acc = fmaf(a[i], b[i], acc);
  }
  return warp_sum(acc);
}

// ------------------------------------------------------------------------------- rerank
// One warp per query row.  (1) pick the kp best GEMM candidates over all gallery ranges,
// (2) re-score them in fp32 from the fp32 masters (same summation order for every candidate, so
// duplicate gallery rows tie exactly), (3) emit the k best ordered (score desc, index asc).
constexpr int kRerankWarps = 4;
constexpr int kMaxKp = 64;

__global__ void __launch_bounds__(kRerankWarps * 32)
rerank_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, long long m,
              long long full_rows, int splits, int kp, int k, const float* __restrict__ q_f32,
              const float* __restrict__ g_f32, int d, float threshold, long long row_offset,
              float* __restrict__ out_sim, long long* __restrict__ out_idx) {
  __shared__ float s_val[kRerankWarps][kMaxKp];
  __shared__ int s_idx[kRerankWarps][kMaxKp];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  if (row < full_rows) splits = 1;     // rows of the unsplit query tiles own one list (SearchPlan)
  const int ncand = splits * kp;
  const size_t cbase = plan_cand_base(full_rows, ncand / kp, kp, row, 0);
  const float* cv = cand_val + cbase;
  const int32_t* ci = cand_idx + cbase;

  int nsel = 0;
  if (splits == 1) {
    for (int e = lane; e < kp; e += 32) {
      s_val[w][e] = cv[e];
      s_idx[w][e] = ci[e];
    }
    nsel = kp;
  } else {
    float bv = INFINITY;
    long long bi = -1;
    for (int t = 0; t < kp; ++t) {
      float v;
      long long i;
      const bool ok = warp_next_best(
          ncand, bv, bi, [&](int e, float& vv, long long& ii) { vv = cv[e]; ii = ci[e]; }, v, i);
      if (!ok) break;
      if (lane == 0) {
        s_val[w][t] = v;
        s_idx[w][t] = static_cast<int>(i);
      }
      bv = v;
      bi = i;
      nsel = t + 1;
    }
    for (int e = nsel + lane; e < kp; e += 32) {
      s_val[w][e] = -INFINITY;
      s_idx[w][e] = -1;
    }
    nsel = kp;
  }
  __syncwarp();

  if (g_f32 != nullptr) {
    const float* q = q_f32 + row * d;
    for (int t = 0; t < nsel; ++t) {
      const int gi = s_idx[w][t];
      if (gi < 0) continue;  // warp-uniform
      const float s = warp_dot_f32(q, g_f32 + static_cast<long long>(gi) * d, d);
      if (lane == 0) s_val[w][t] = s;
    }
    __syncwarp();
  }

  float bv = INFINITY;
  long long bi = -1;
  for (int j = 0; j < k; ++j) {
    float v = -INFINITY;
    long long i = -1;
    bool ok = false;
    if (bi != LLONG_MIN) {
      ok = warp_next_best(
          nsel, bv, bi,
          [&](int e, float& vv, long long& ii) { vv = s_val[w][e]; ii = s_idx[w][e]; }, v, i);
    }
    if (ok && v >= threshold) {
      bv = v;
      bi = i;
      if (lane == 0) {
        out_sim[row * k + j] = v;
        out_idx[row * k + j] = i + row_offset;
      }
    } else {
      bi = LLONG_MIN;  // exhausted (or below threshold: everything after is too)
      if (lane == 0) {
        out_sim[row * k + j] = -INFINITY;
        out_idx[row * k + j] = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------- sharded search
// Phase 1 on every shard: best kp GEMM candidates of each query row over the shard's ranges, written
// with GLOBAL indices either to a local [m, kp] list or straight into the receive buffer of the rank
// that owns the row's query slice (peer HBM mapped through CUDA IPC: the candidate exchange is this
// kernel's store stream over NVLink, there is no collective).
__global__ void __launch_bounds__(kRerankWarps * 32)
select_candidates_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx,
                         long long m, long long full_rows, int splits, int kp, long long row_offset, const ScatterSpec sc,
                         float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  const size_t cbase = plan_cand_base(full_rows, splits, kp, row, 0);
  if (row < full_rows) splits = 1;
  const int ncand = splits * kp;
  const float* cv = cand_val + cbase;
  const int32_t* ci = cand_idx + cbase;
  float* dv;
  long long* di;
  if (sc.n_slices > 0) {
    const long long j = row / sc.rows_per_slice, lr = row - j * sc.rows_per_slice;
    const long long rows_j = min(sc.rows_per_slice, m - j * sc.rows_per_slice);
    const long long o = (static_cast<long long>(sc.slot) * rows_j + lr) * kp;
    dv = sc.val[j] + o;
    di = sc.idx[j] + o;
  } else {
    dv = out_val + row * kp;
    di = out_idx + row * kp;
  }
  if (splits == 1) {
    for (int e = lane; e < kp; e += 32) {
      const int i = ci[e];
      dv[e] = i >= 0 ? cv[e] : -INFINITY;
      di[e] = i >= 0 ? i + row_offset : -1;
    }
    return;
  }
  // lane t keeps the t-th best (and t + 32 for kp = 64) and the row leaves as two coalesced stores:
  // single-lane stores would cross NVLink as 2 * kp tiny writes per row
  float bv = INFINITY;
  long long bi = -1;
  float keep_v[2] = {-INFINITY, -INFINITY};
  long long keep_i[2] = {-1, -1};
  for (int t = 0; t < kp; ++t) {
    float v;
    long long i;
    const bool ok = warp_next_best(
        ncand, bv, bi, [&](int e, float& vv, long long& ii) { vv = cv[e]; ii = ci[e]; }, v, i);
    if (!ok) break;
    if (lane == (t & 31)) {
      keep_v[t >> 5] = v;
      keep_i[t >> 5] = i + row_offset;
    }
    bv = v;
    bi = i;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = lane + 32 * h;
    if (e < kp) {
      dv[e] = keep_v[h];
      di[e] = keep_i[h];
    }
  }
}

// Phase 2 on the owner of a query slice: the `parts` lists of a row ([parts, m, kp], global indices)
// are merged to the kp best by GEMM score, re-scored in fp32 from the masters of the owning shards
// (own HBM or peer HBM over NVLink) and the k best leave ordered (score desc, index asc) - the same
// selection the single-GPU rerank makes over its gallery ranges.
__global__ void __launch_bounds__(kRerankWarps * 32)
rerank_merged_kernel(const float* __restrict__ cand_val, const long long* __restrict__ cand_idx,
                     long long m, int parts, int kp, int k, const float* __restrict__ q_f32,
                     const RowSource src, int d, float threshold, float* __restrict__ out_sim,
                     long long* __restrict__ out_idx) {
  __shared__ float s_val[kRerankWarps][kMaxKp];
  __shared__ long long s_idx[kRerankWarps][kMaxKp];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  const int ncand = parts * kp;
  auto fetch = [&](int e, float& vv, long long& ii) {
    const int pt = e / kp, t = e - pt * kp;
    const long long o = (static_cast<long long>(pt) * m + row) * kp + t;
    vv = cand_val[o];
    ii = cand_idx[o];
  };
  {
    float bv = INFINITY;
    long long bi = -1;
    int t = 0;
    for (; t < kp; ++t) {
      float v;
      long long i;
      if (!warp_next_best(ncand, bv, bi, fetch, v, i)) break;
      if (lane == 0) {
        s_val[w][t] = v;
        s_idx[w][t] = i;
      }
      bv = v;
      bi = i;
    }
    for (int e = t + lane; e < kp; e += 32) {
      s_val[w][e] = -INFINITY;
      s_idx[w][e] = -1;
    }
  }
  __syncwarp();
  const float* q = q_f32 + row * d;
  for (int t = 0; t < kp; ++t) {
    const long long gi = s_idx[w][t];
    if (gi < 0) continue;  // warp-uniform
    int part = -1;
    for (int pp = 0; pp < src.nparts; ++pp)
      if (gi >= src.off[pp] && gi < src.off[pp] + src.n[pp]) part = pp;
    if (part < 0 || src.f32[part] == nullptr) {
      if (lane == 0) s_idx[w][t] = -1;
      continue;
    }
    const float s = warp_dot_f32(q, src.f32[part] + (gi - src.off[part]) * d, d);
    if (lane == 0) s_val[w][t] = s;
  }
  __syncwarp();
  float bv = INFINITY;
  long long bi = -1;
  for (int j = 0; j < k; ++j) {
    float v = -INFINITY;
    long long i = -1;
    bool ok = false;
    if (bi != LLONG_MIN)
      ok = warp_next_best(
          kp, bv, bi, [&](int e, float& vv, long long& ii) { vv = s_val[w][e]; ii = s_idx[w][e]; }, v, i);
    if (ok && v >= threshold) {
      bv = v;
      bi = i;
      if (lane == 0) {
        out_sim[row * k + j] = v;
        out_idx[row * k + j] = i;
      }
    } else {
      bi = LLONG_MIN;
      if (lane == 0) {
        out_sim[row * k + j] = -INFINITY;
        out_idx[row * k + j] = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------- sharded search, re-score at the shards
// Phase 2 done where the rows live (round 2).  rerank_merged_kernel PULLS the KP fp32 master rows of every query
// row from the shards that own them - 16 x 3 KB per row over NVLink, 0.44 GB per rank and search at 8 GPUs, the
// largest non-GEMM item of the step.  Here the traffic goes the other way and is ~30x smaller:
//   (1) the owner of a query slice merges the P candidate lists of each of its rows to the KP best by GEMM
//       score and stores that index list (KP x 8 bytes) into every shard's request area       [exchange_merge]
//   (2) every shard walks the request lists of ALL rows, re-scores the entries that fall into its own row
//       range in fp32 - from its local master and the fp32 query rows, which every rank received by copy
//       engine under the GEMM - and stores each score into the owner's score area                [exchange_rescore]
//   (3) the owner orders (fp32 score desc, index asc), applies the threshold, emits the top-k     [exchange_finalize]
// with one stream-ordered barrier between the steps.  Same candidates, same warp_dot_f32 on the same operands
// as the single-GPU rerank_kernel, so the results stay bit-identical to the unsharded search.
__global__ void __launch_bounds__(kRerankWarps * 32)
exchange_merge_kernel(const float* __restrict__ cand_val, const long long* __restrict__ cand_idx, long long m,
                      int parts, int kp, const ReqDst dst) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  const int ncand = parts * kp;
  auto fetch = [&](int e, float& vv, long long& ii) {
    const int pt = e / kp, t = e - pt * kp;
    const long long o = (static_cast<long long>(pt) * m + row) * kp + t;
    vv = cand_val[o];
    ii = cand_idx[o];
  };
  // lane t keeps the t-th best (and t + 32 for kp = 64): the list leaves as coalesced stores
  float bv = INFINITY;
  long long bi = -1;
  long long keep[2] = {-1, -1};
  for (int t = 0; t < kp; ++t) {
    float v;
    long long i;
    if (!warp_next_best(ncand, bv, bi, fetch, v, i)) break;
    if (lane == (t & 31)) keep[t >> 5] = i;
    bv = v;
    bi = i;
  }
  for (int s = 0; s < dst.n; ++s) {
    long long* o = dst.req[s] + row * kp;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = lane + 32 * h;
      if (e < kp) o[e] = keep[h];
    }
  }
}

__global__ void __launch_bounds__(kRerankWarps * 32)
exchange_rescore_kernel(const long long* __restrict__ req, const float* __restrict__ q_f32,
                        const float* __restrict__ g_f32, long long g_off, long long g_n, int d, int owners,
                        long long rows_per_slice, long long m_total, int kp, const ScoreDst dst) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;   // global query row
  if (row >= m_total) return;
  const int o = static_cast<int>(row / rows_per_slice);
  const long long lr = row - static_cast<long long>(o) * rows_per_slice;
  if (o >= owners) return;
  // request area: [owner][rows_per_slice][kp]
  const long long* lst = req + (static_cast<long long>(o) * rows_per_slice + lr) * kp;
  float* out = dst.score[o] + lr * kp;
  const float* q = q_f32 + row * d;
  for (int h = 0; h * 32 < kp; ++h) {
    const int e = lane + 32 * h;
    const long long gi = e < kp ? lst[e] : -1;
    const bool mine = gi >= g_off && gi < g_off + g_n;
    unsigned todo = __ballot_sync(kFull, mine);
    float s_mine = 0.f;
    while (todo) {
      const int t = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long g = __shfl_sync(kFull, gi, t);
      const float s = warp_dot_f32(q, g_f32 + (g - g_off) * d, d);
      if (lane == t) s_mine = s;
    }
    if (mine) out[e] = s_mine;
  }
}

__global__ void __launch_bounds__(kRerankWarps * 32)
exchange_finalize_kernel(const long long* __restrict__ req, const float* __restrict__ score, long long m, int kp,
                         int k, float threshold, float* __restrict__ out_sim, long long* __restrict__ out_idx) {
  __shared__ float s_val[kRerankWarps][kMaxKp];
  __shared__ long long s_idx[kRerankWarps][kMaxKp];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  for (int e = lane; e < kp; e += 32) {
    const long long gi = req[row * kp + e];
    s_idx[w][e] = gi;
    s_val[w][e] = gi >= 0 ? score[row * kp + e] : -INFINITY;
  }
  __syncwarp();
  float bv = INFINITY;
  long long bi = -1;
  for (int j = 0; j < k; ++j) {
    float v = -INFINITY;
    long long i = -1;
    bool ok = false;
    if (bi != LLONG_MIN)
      ok = warp_next_best(
          kp, bv, bi, [&](int e, float& vv, long long& ii) { vv = s_val[w][e]; ii = s_idx[w][e]; }, v, i);
    if (ok && v >= threshold) {
      bv = v;
      bi = i;
      if (lane == 0) {
        out_sim[row * k + j] = v;
        out_idx[row * k + j] = i;
      }
    } else {
      bi = LLONG_MIN;
      if (lane == 0) {
        out_sim[row * k + j] = -INFINITY;
        out_idx[row * k + j] = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------- merge_topk
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ in_sim, const long long* __restrict__ in_idx, long long m,
                  int parts, int k, float* __restrict__ out_sim, long long* __restrict__ out_idx) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + w;
  if (row >= m) return;
  const int n = parts * k;
  const float* sv = in_sim + row * n;
  const long long* si = in_idx + row * n;
  float bv = INFINITY;
  long long bi = -1;
  bool done = false;
  for (int j = 0; j < k; ++j) {
    float v = -INFINITY;
    long long i = -1;
    bool ok = false;
    if (!done)
      ok = warp_next_best(
          n, bv, bi, [&](int e, float& vv, long long& ii) { vv = sv[e]; ii = si[e]; }, v, i);
    if (ok) {
      bv = v;
      bi = i;
    } else {
      done = true;
      v = -INFINITY;
      i = -1;
    }
    if (lane == 0) {
      out_sim[row * k + j] = v;
      out_idx[row * k + j] = i;
    }
  }
}

// ------------------------------------------------------------------------------- k-occurrence
// Kernel (c): N_k(j) = #{rows i : j in topk(i)} (references/Adversarial_Hubness_.../README.md:43-57).
// The int64 index stream is read with 128-bit loads at HBM speed; what bounds the kernel is the
// atomic path, and on the SM side: a RED costs the load/store pipe ~1.3 cycles per LANE whatever the
// address (B300_MICROARCH "REDG 1.29 cyc/lane spread"; measured here ~180 G increments/s over 148 SMs), a
// shared-memory atomic 1-2 cycles per lane, so no private copy makes an increment cheaper than its RED.
// What the kernel can do is (1) not serialise: increments that land in ONE 128-byte line queue in its L2
// slice at ~1.5 G/s, and hubness histograms are exactly the data with hot lines (an adversarial hub sits
// in most rows; popular rows cluster) - and (2) send fewer REDs where the stream repeats itself.  So:
//   * histograms that fit in shared memory (<= 48 KB) and see many increments per bin accumulate
//     privately per block with shared-memory atomics and flush once;
//   * otherwise a sampling pre-pass (one block, <= 8192 strided entries) finds the 32-bin lines that
//     hold >= 1/512 of the stream - the share above which a line's serial time in its slice would reach the
//     kernel's own - and publishes up to 256 of them; the main kernel keeps those lines (32 counters
//     each) in a block-private shared-memory table (one LDS probe per entry, shared-memory atomic on a
//     hit, flushed once per block) and sends every other entry to its bin with one RED;
//   * WARP-AGGREGATED atomics (north_star (c)): the lanes of a warp that hold the same bin elect one lane,
//     which adds the group's population (__match_any_sync).  A warp-wide match costs the SM more than the
//     32 REDs it can save (measured: 50 M i.i.d. entries 420 -> 950 us with the vote on every entry), so
//     the pre-pass also MEASURES how often the kernel's own lane groups repeat a bin (32 windows of the
//     stream, same entry-to-lane mapping) and publishes that rate; the main kernel votes only when at
//     least kAggPermille of the entries would be absorbed (top-k lists of neighbouring rows - the 5 text
//     variants of one query - and hub-dominated streams are; a shuffled stream is not).
constexpr int kHotMax = 256;          // hot lines the sampling pass may publish
constexpr int kHotSlots = 1024;       // open-addressing slots of the per-block key table
constexpr int kSampleSlots = 4096;
constexpr int kSamples = 8192;
constexpr int kHotShare = 512;        // hot = sample count * kHotShare >= samples
constexpr int kAggSlot = kHotMax + 1; // scratch word: permille of entries a warp vote would absorb
constexpr int kAggPermille = 500;     // vote when the average group holds >= 2 lanes

__device__ __forceinline__ unsigned bin_hash(int b) {
  return static_cast<unsigned>(b) * 0x9E3779B1u;
}

// scratch layout: hot[0] = number of keys, hot[1 .. kHotMax] = line numbers ((bin - idx_base) >> 5),
// hot[kAggSlot] = permille of the sampled entries that share their bin with a lower lane of their group
__global__ void __launch_bounds__(1024)
k_occurrence_sample_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                           long long n_bins, int* __restrict__ hot) {
  __shared__ int s_cnt[kSampleSlots];     // lossy hashed counts
  __shared__ int s_key[kHotSlots];        // exact table of the candidates
  __shared__ int s_exact[kHotSlots];
  __shared__ int s_n, s_dup, s_valid;
  for (int i = threadIdx.x; i < kSampleSlots; i += blockDim.x) s_cnt[i] = 0;
  for (int i = threadIdx.x; i < kHotSlots; i += blockDim.x) {
    s_key[i] = -1;
    s_exact[i] = 0;
  }
  if (threadIdx.x == 0) s_n = s_dup = s_valid = 0;
  __syncthreads();
  const int samples = total < kSamples ? static_cast<int>(total) : kSamples;
  const long long stride = total / samples;
  const int thresh = max(4, samples / kHotShare);
#pragma unroll 8
  for (int i = threadIdx.x; i < samples; i += blockDim.x) {
    const long long b = __ldg(idx + static_cast<long long>(i) * stride) - idx_base;
    if (b >= 0 && b < n_bins) atomicAdd(&s_cnt[bin_hash(static_cast<int>(b >> 5)) >> 20], 1);
  }
  __syncthreads();
  // second walk: entries whose lossy slot is heavy are counted exactly (bounded linear probing; a
  // dropped candidate only costs speed, never correctness)
#pragma unroll 4
  for (int i = threadIdx.x; i < samples; i += blockDim.x) {
    const long long b64 = __ldg(idx + static_cast<long long>(i) * stride) - idx_base;
    if (b64 < 0 || b64 >= n_bins) continue;
    const int line = static_cast<int>(b64 >> 5);
    if (s_cnt[bin_hash(line) >> 20] < thresh) continue;
    unsigned h = (bin_hash(line) >> 8) & (kHotSlots - 1);
    for (int probe = 0; probe < 16; ++probe) {
      const int k = atomicCAS(&s_key[h], -1, line);
      if (k == -1 || k == line) {
        atomicAdd(&s_exact[h], 1);
        break;
      }
      h = (h + 1) & (kHotSlots - 1);
    }
  }
  __syncthreads();
  // repetition inside the main kernel's lane groups, hot lines aside (they never reach the vote): warp w reads
  // the 128 consecutive entries at w/32 of the stream the way a trip of the main kernel does (lane l holds
  // entries 4l .. 4l+3, group h = the lanes' h-th entries)
  {
    __syncwarp();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long base = ((total / 32) * w) & ~3ll;
    int dup = 0, valid = 0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const long long e = base + 4 * lane + h;
      long long b = e < total ? __ldg(idx + e) - idx_base : -1;
      if (b >= n_bins) b = -1;
      if (b >= 0) {
        const int line = static_cast<int>(b >> 5);
        unsigned hh = (bin_hash(line) >> 8) & (kHotSlots - 1);
        for (int probe = 0; probe < 16; ++probe) {
          const int k = s_key[hh];
          if (k == -1) break;
          if (k == line) {
            if (s_exact[hh] >= thresh) b = -1;
            break;
          }
          hh = (hh + 1) & (kHotSlots - 1);
        }
      }
      __syncwarp();
      const long long key = b >= 0 ? b : -1ll - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (b >= 0) {
        ++valid;
        if (lane != __ffs(peers) - 1) ++dup;
      }
    }
    if (valid) atomicAdd(&s_valid, valid);
    if (dup) atomicAdd(&s_dup, dup);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHotSlots; i += blockDim.x)
    if (s_key[i] >= 0 && s_exact[i] >= thresh) {
      const int slot = atomicAdd(&s_n, 1);
      if (slot < kHotMax) hot[1 + slot] = s_key[i];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    hot[0] = min(s_n, kHotMax);
    hot[kAggSlot] = s_valid > 0 ? static_cast<int>(1000ll * s_dup / s_valid) : 0;
  }
}

// Walks the stream in trips of 8 entries per thread (four 128-bit loads in flight when VEC), then 4, then 1;
// `emit(b, n)` receives the trip's n bins at once, -1 for an entry outside the histogram, so that the callee
// can issue its n table probes / votes / REDs back to back instead of one dependent chain per entry (ncu on
// the per-entry form: 52 instructions per entry, 28 of every 50 stall cycles on the shared-memory scoreboard).
template <bool VEC, typename Emit>
__device__ __forceinline__ void for_each_bin(const long long* __restrict__ idx, long long total,
                                             long long idx_base, long long n_bins, Emit emit) {
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long quads = VEC ? (total >> 2) : 0;
  const longlong2* idx2 = reinterpret_cast<const longlong2*>(idx);
  auto bin_of = [&](long long v) {
    v -= idx_base;
    return (v >= 0 && v < n_bins) ? v : -1ll;
  };
  long long p = tid;
  for (; p + nthreads < quads; p += 2 * nthreads) {
    const longlong2 a0 = idx2[2 * p], c0 = idx2[2 * p + 1];
    const longlong2 a1 = idx2[2 * (p + nthreads)], c1 = idx2[2 * (p + nthreads) + 1];
    const long long b[8] = {bin_of(a0.x), bin_of(a0.y), bin_of(c0.x), bin_of(c0.y),
                            bin_of(a1.x), bin_of(a1.y), bin_of(c1.x), bin_of(c1.y)};
    emit(b, 8);
  }
  // (p + nthreads < quads is not warp-uniform on the last trip: the leftovers run under the lanes' own mask;
  // red_aggregated votes over __activemask())
  for (; p < quads; p += nthreads) {
    const longlong2 a = idx2[2 * p], c = idx2[2 * p + 1];
    const long long b[8] = {bin_of(a.x), bin_of(a.y), bin_of(c.x), bin_of(c.y), -1, -1, -1, -1};
    emit(b, 4);
  }
  for (long long e = (quads << 2) + tid; e < total; e += nthreads) {
    const long long b[8] = {bin_of(idx[e]), -1, -1, -1, -1, -1, -1, -1};
    emit(b, 1);
  }
}

// One RED per DISTINCT bin among the lanes of a warp that reach this point together (what north_star calls
// warp-aggregated atomics): the lanes holding the same bin elect their lowest lane, which adds the group's
// population.  Lanes without a bin (b < 0) take part in the vote with a value nobody shares.  NARROW: the
// bins fit 31 bits, the vote compares 32-bit keys (measured: the 64-bit match costs 2.2x the kernel, the
// 32-bit one 6 %).
template <bool NARROW>
__device__ __forceinline__ void red_aggregated(int* __restrict__ counts, long long b) {
  const unsigned active = __activemask();
  const int lane = threadIdx.x & 31;
  unsigned peers;
  if (NARROW) {
    peers = __match_any_sync(active, b >= 0 ? static_cast<int>(b) : -1 - lane);
  } else {
    peers = __match_any_sync(active, b >= 0 ? b : -1ll - lane);
  }
  if (b >= 0 && lane == __ffs(peers) - 1) atomicAdd(&counts[b], __popc(peers));
}

template <bool VEC>
__global__ void __launch_bounds__(256)
k_occurrence_smem_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                         long long n_bins, int* __restrict__ counts) {
  extern __shared__ int s_hist[];
  for (int b = threadIdx.x; b < n_bins; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  for_each_bin<VEC>(idx, total, idx_base, n_bins, [&](const long long (&b)[8], int n) {
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < n && b[h] >= 0) atomicAdd(&s_hist[b[h]], 1);
  });
  __syncthreads();
  for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
    const int c = s_hist[b];
    if (c) atomicAdd(&counts[b], c);
  }
}

constexpr unsigned kSlotEmpty = 0xffffffffu;   // key table word: (line << 8) | compact line id, lines < 2^24

template <bool VEC>
__global__ void __launch_bounds__(256)
k_occurrence_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                    long long n_bins, int* __restrict__ counts, const int* __restrict__ hot, int force_agg) {
  __shared__ unsigned s_tab[kHotSlots];
  __shared__ int s_cnt[kHotMax * 32];
  // block-uniform (every block reads the same scratch words)
  const int n_hot = hot ? min(hot[0], kHotMax) : 0;
  const bool agg = force_agg >= 0 ? force_agg != 0 : (hot != nullptr && hot[kAggSlot] >= kAggPermille);
  if (n_hot > 0) {
    for (int i = threadIdx.x; i < kHotSlots; i += blockDim.x) s_tab[i] = kSlotEmpty;
    for (int i = threadIdx.x; i < n_hot * 32; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_hot; i += blockDim.x) {
      const int line = hot[1 + i];
      unsigned h = (bin_hash(line) >> 8) & (kHotSlots - 1);
      const unsigned word = (static_cast<unsigned>(line) << 8) | static_cast<unsigned>(i);
      while (atomicCAS(&s_tab[h], kSlotEmpty, word) != kSlotEmpty) h = (h + 1) & (kHotSlots - 1);   // keys are distinct
    }
    __syncthreads();
  }
  // One trip: all first probes are issued together, then each entry is resolved - an empty slot (the common
  // case: the table is at most a quarter full) sends it to its bin, a key match to the block's private counters,
  // anything else walks on.
  auto trip = [&](const long long (&b)[8], int n, auto cold) {
    unsigned slot[8], word[8];
    if (n_hot > 0) {
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        slot[h] = (bin_hash(static_cast<int>(b[h] >> 5)) >> 8) & (kHotSlots - 1);
        word[h] = h < n ? s_tab[slot[h]] : kSlotEmpty;
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h >= n) break;
      long long bb = b[h];
      if (n_hot > 0 && bb >= 0 && word[h] != kSlotEmpty) {
        const unsigned line = static_cast<unsigned>(bb >> 5);
        unsigned w = word[h], s = slot[h];
        while (w != kSlotEmpty && (w >> 8) != line) {
          s = (s + 1) & (kHotSlots - 1);
          w = s_tab[s];
        }
        if (w != kSlotEmpty) {
          atomicAdd(&s_cnt[(w & 255u) * 32 + (static_cast<int>(bb) & 31)], 1);
          bb = -1;
        }
      }
      cold(bb);
    }
  };
  if (!agg) {
    for_each_bin<VEC>(idx, total, idx_base, n_bins, [&](const long long (&b)[8], int n) {
      trip(b, n, [&](long long bb) {
        if (bb >= 0) atomicAdd(&counts[bb], 1);
      });
    });
  } else if (n_bins <= 0x7fffffffll) {
    for_each_bin<VEC>(idx, total, idx_base, n_bins, [&](const long long (&b)[8], int n) {
      trip(b, n, [&](long long bb) { red_aggregated<true>(counts, bb); });   // hot / invalid lanes vote with no bin
    });
  } else {
    for_each_bin<VEC>(idx, total, idx_base, n_bins, [&](const long long (&b)[8], int n) {
      trip(b, n, [&](long long bb) { red_aggregated<false>(counts, bb); });
    });
  }
  if (n_hot > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_hot * 32; i += blockDim.x) {
      const int c = s_cnt[i];
      if (c) atomicAdd(&counts[static_cast<long long>(hot[1 + (i >> 5)]) * 32 + (i & 31)], c);
    }
  }
}

// ---- bucketed path: streams of millions of entries over histograms far beyond shared memory ----------------
// Every entry of the RED path costs one atomic in the L2 (~180 G/s chip-wide, a fifth of what HBM delivers as
// index stream).  The only way past that is to make the increments shared-memory ones, i.e. to bring the entries
// of one bin range together first:
//   pass 1 (k_occurrence_partition_tma_kernel / k_occurrence_partition_kernel): a CTA takes a TILE of 8192 entries
//     and writes it back bucket-sorted
//     (bucket = bin >> 15) as 16-bit keys (bin & 32767) - tile-major, so the write is one contiguous 16 KB block
//     and no bucket can overflow whatever the distribution - plus the tile's bucket boundaries (transposed,
//     offs[bucket][tile]) and the bucket totals.  The rank of an entry inside its (thread, bucket) cell comes from
//     THREAD-PRIVATE 16-bit counters (s_cell[bucket][thread]: plain conflict-free LDS + STS, 1/32 cycle per lane
//     where a shared-memory atomic with return costs ~2), the cell bases from one in-place scan per bucket.
//   pass 2 (k_occurrence_bucket_kernel): CTAs are dealt to the buckets in proportion to their totals (every CTA
//     derives the same deal from the totals; hub-heavy buckets get many), a CTA counts its share of a bucket's
//     segments - one segment per tile, the warps taking tiles in turn - in a 128 KB shared-memory histogram and
//     flushes the non-zero counters with REDs to consecutive addresses.
// HBM traffic: 8 B/entry read + 2 B written + 2 B read (+ 4 B/bin) against 8 B/entry of the single pass.
// Measured on a B200 (profiles/r4i_probe.log, r4j_probe.log, r4k_probe.log; 1 M bins, stream idx = N u^3): 50 M
// entries 200 us against 456 us (0.31 against 0.14 of the HBM peak; uniform stream 201 against 336), 5 M entries 50
// against 84, 1 M entries 37 against 45.  Pass 1 is bound by the load/store pipe (ncu: 72 % busy, 3 900 shared-memory
// wavefronts per tile of which ~1 000 are bank conflicts of the 16-bit scatter), pass 2 by the shared-memory atomics.
constexpr int kPartThreads = 512;
constexpr int kPartPer = 16;                               // entries per thread and tile
constexpr int kPartTile = kPartThreads * kPartPer;         // 8192
constexpr int kPartShift = 15;
constexpr int kPartWidth = 1 << kPartShift;                // bins per bucket: 128 KB of counters
constexpr int kPartMaxBuckets = 127;                       // histograms up to 4 161 536 bins (bucket 127 = no bin)
constexpr int kPartTotalStride = 16;                       // one 64-bit bucket total per 128-byte line
constexpr int kCountThreads = 1024;

// global -> shared bulk copy (TMA, no tensor map), completion bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s_aux(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += y;
  }
  return v;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kPartThreads, MIN_BLOCKS)
k_occurrence_partition_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                              long long n_bins, int n_buckets, long long n_tiles,
                              unsigned short* __restrict__ keys, unsigned short* __restrict__ offs,
                              unsigned long long* __restrict__ totals, long long tile0) {
  // s_cell[bucket][thread]: first the number of entries thread `thread` holds for `bucket`, then (in place) the
  // position of its first one inside the bucket.  16-bit cells: lanes 2i and 2i+1 share a bank (a 2-way conflict
  // when they address different buckets) but the address is one multiply-add - the kernel is bound by its
  // instruction count (ncu on the first version: 75 instructions per entry, issue slots 62 % busy, HBM 37 %), not
  // by shared memory.  Bucket n_buckets is the bin-less one (unused slots, rows of other shards, the ragged end of
  // the stream): every entry takes the same straight-line path and the bin-less ones land behind the tile's keys.
  extern __shared__ __align__(16) unsigned char part_smem[];
  unsigned short* s_cell = reinterpret_cast<unsigned short*>(part_smem);
  unsigned short* s_sorted = s_cell + static_cast<size_t>(n_buckets + 1) * kPartThreads;
  __shared__ int s_tot[kPartMaxBuckets + 1];
  __shared__ int s_base[kPartMaxBuckets + 2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // cell of thread (warp w, lane l): 16-bit half (w & 1) of word (w >> 1) * 32 + l - the lanes of a warp sit in 32
  // different banks whatever buckets they address
  const int cell = ((warp >> 1) << 6) | (lane << 1) | (warp & 1);
  const long long tile = tile0 + blockIdx.x;
  const long long e0 = tile * kPartTile;
  const longlong2* idx2 = reinterpret_cast<const longlong2*>(idx);
  const bool full = e0 + kPartTile <= total;
  const unsigned nb = static_cast<unsigned>(n_bins);                       // <= 127 * 32768 on this path
  const unsigned binless = static_cast<unsigned>(n_buckets) << kPartShift;

  auto load_half = [&](int half, long long (&v)[kPartPer / 2]) {
#pragma unroll
    for (int j = 0; j < kPartPer / 4; ++j) {
      const long long e = e0 + static_cast<long long>((half * (kPartPer / 4) + j) * kPartThreads + tid) * 2;
      long long a = idx_base - 1, c = idx_base - 1;          // outside the histogram
      if (full || e + 1 < total) {
        const longlong2 t = __ldcs(idx2 + (e >> 1));
        a = t.x;
        c = t.y;
      } else if (e < total) {
        a = idx[e];
      }
      v[2 * j] = a;
      v[2 * j + 1] = c;
    }
  };
  unsigned ent[kPartPer];          // bin (22 bits: bucket << 15 | key) | rank inside the (thread, bucket) cell << 22
  auto count_half = [&](int half, const long long (&v)[kPartPer / 2]) {
#pragma unroll
    for (int j = 0; j < kPartPer / 2; ++j) {
      const long long b = v[j] - idx_base;
      const unsigned lo = static_cast<unsigned>(b), hi = static_cast<unsigned>(static_cast<unsigned long long>(b) >> 32);
      const unsigned sel = (hi == 0u && lo < nb) ? lo : binless;
      unsigned short* c = s_cell + (sel >> kPartShift) * kPartThreads + cell;
      const unsigned old = *c;
      *c = static_cast<unsigned short>(old + 1u);
      ent[half * (kPartPer / 2) + j] = sel | (old << 22);
    }
  };
  long long v[kPartPer / 2];
  load_half(0, v);
  {
    uint4* z = reinterpret_cast<uint4*>(part_smem);
    const int n16 = (n_buckets + 1) * kPartThreads * 2 / 16;
    for (int i = tid; i < n16; i += kPartThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  count_half(0, v);
  load_half(1, v);
  count_half(1, v);
  __syncthreads();
  // cell bases, in place: a warp takes a bucket; the order of the cells in memory is the order of the threads inside
  // the bucket (any fixed order does)
  for (int b = warp; b <= n_buckets; b += kPartThreads / 32) {
    unsigned* c = reinterpret_cast<unsigned*>(s_cell + b * kPartThreads);
    unsigned w[kPartThreads / 64];
    unsigned run = 0;
#pragma unroll
    for (int j = 0; j < kPartThreads / 64; ++j) w[j] = c[j * 32 + lane];
#pragma unroll
    for (int j = 0; j < kPartThreads / 64; ++j) {
      const unsigned lo = w[j] & 0xffffu, hi = w[j] >> 16;
      w[j] = run | ((run + lo) << 16);
      run += lo + hi;
    }
    const unsigned incl = warp_inclusive_scan(run, lane);
    const unsigned base2 = (incl - run) * 0x10001u;          // a tile holds 8192 entries: the halves never carry
#pragma unroll
    for (int j = 0; j < kPartThreads / 64; ++j) c[j * 32 + lane] = w[j] + base2;
    if (lane == 31) s_tot[b] = static_cast<int>(incl);
  }
  __syncthreads();
  if (warp == 0) {
    unsigned t[4], run = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = lane * 4 + i;
      t[i] = b <= n_buckets ? static_cast<unsigned>(s_tot[b]) : 0u;
      run += t[i];
    }
    const unsigned incl = warp_inclusive_scan(run, lane);
    unsigned base = incl - run;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = lane * 4 + i;
      if (b <= n_buckets) s_base[b] = static_cast<int>(base);
      base += t[i];
    }
  }
  __syncthreads();
  // bucket boundaries of the tile (s_base[n_buckets] = where the bin-less entries start = the number of keys)
  if (tid <= n_buckets) offs[static_cast<long long>(tid) * n_tiles + tile] = static_cast<unsigned short>(s_base[tid]);
  if (tid < n_buckets && s_tot[tid] > 0)
    atomicAdd(totals + tid * kPartTotalStride, static_cast<unsigned long long>(s_tot[tid]));
#pragma unroll
  for (int j = 0; j < kPartPer; ++j) {
    const unsigned w = ent[j];
    const unsigned bucket = (w >> kPartShift) & 127u;
    const unsigned pos = static_cast<unsigned>(s_base[bucket]) + s_cell[bucket * kPartThreads + cell] + (w >> 22);
    s_sorted[pos] = static_cast<unsigned short>(w & (kPartWidth - 1));
  }
  __syncthreads();
  const int n16 = (s_base[n_buckets] * 2 + 15) >> 4;
  uint4* dst = reinterpret_cast<uint4*>(keys + tile * kPartTile);
  const uint4* src = reinterpret_cast<const uint4*>(s_sorted);
  for (int i = tid; i < n16; i += kPartThreads) __stcs(dst + i, src[i]);
}

// The same pass as a persistent kernel fed by the copy engine (histograms up to 31 buckets = 1 015 808 bins, two
// CTAs per SM): ncu on the kernel above shows neither pipe saturated (LSU ~56 %, issue ~55 %) - the CTAs spend
// ~40 % of their time waiting for their own index loads, and three 512-thread CTAs per SM do not overlap that away.
// Here one thread hands the next tile (64 KB of int64 indices) to cp.async.bulk as soon as the count phase has
// consumed the current one, so the load runs under the scan / scatter / copy-out phases; the count phase reads
// the indices with conflict-free 128-bit LDS.  Same cells, same order, same output as the kernel above.
// Two other ways of ranking were measured on the 50 M-entry stream and dropped (profiles/r4f_probe.log,
// r4g_probe.log): five ballots over the bits of the bucket number with the running counts in registers (no shared
// memory while counting, but 67 integer instructions per entry: ALU pipe 84 % busy, 158 us against 145) and
// MATCH.ANY on the bucket number with one running count per (warp, bucket) (200 us skewed / 251 us uniform - the
// match instruction takes time in proportion to the number of distinct values it finds).
constexpr int kTmaCellBuckets = 32;                       // rows of cells: 31 buckets + the bin-less one
constexpr size_t kTmaRawBytes = static_cast<size_t>(kPartTile) * 8;
constexpr size_t kTmaSmemBytes = kTmaRawBytes + static_cast<size_t>(kTmaCellBuckets) * kPartThreads * 2 +
                                 static_cast<size_t>(kPartTile) * 2 + 2 * (kTmaCellBuckets + 1) * 4 + 16;

__global__ void __launch_bounds__(kPartThreads, 2)
k_occurrence_partition_tma_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                                  long long n_bins, int n_buckets, long long n_tiles,
                                  unsigned short* __restrict__ keys, unsigned short* __restrict__ offs,
                                  unsigned long long* __restrict__ totals) {
  extern __shared__ __align__(128) unsigned char part_smem[];
  const longlong2* s_raw = reinterpret_cast<const longlong2*>(part_smem);
  unsigned short* s_cell = reinterpret_cast<unsigned short*>(part_smem + kTmaRawBytes);
  unsigned short* s_sorted = s_cell + kTmaCellBuckets * kPartThreads;
  int* s_tot = reinterpret_cast<int*>(s_sorted + kPartTile);
  int* s_base = s_tot + kTmaCellBuckets + 1;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_base + kTmaCellBuckets + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cell = ((warp >> 1) << 6) | (lane << 1) | (warp & 1);
  const unsigned nb = static_cast<unsigned>(n_bins);
  const unsigned binless = static_cast<unsigned>(n_buckets) << kPartShift;
  const int n_cell16 = (n_buckets + 1) * kPartThreads * 2 / 16;
  long long tile = blockIdx.x;
  if (tile >= n_tiles) return;
  // Hands tile t to the copy engine (called by every thread, after a barrier that ends all reads of the previous
  // tile).  The ragged end of the stream: the copy takes the 16-byte multiple, the threads write the odd entry and
  // fill the empty slots with an index outside the histogram.
  auto issue = [&](long long t) {
    const long long first = t * kPartTile, left = total - first;
    const unsigned n_ent = left < kPartTile ? static_cast<unsigned>(left) : static_cast<unsigned>(kPartTile);
    const unsigned bytes = (n_ent * 8u) & ~15u;
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_expect_tx(s_full, bytes);
      if (bytes) bulk_g2s_aux(part_smem, idx + first, bytes, s_full);
    }
    if (n_ent < static_cast<unsigned>(kPartTile)) {
      long long* raw = reinterpret_cast<long long*>(part_smem);
      for (unsigned i = (bytes >> 3) + tid; i < static_cast<unsigned>(kPartTile); i += kPartThreads)
        raw[i] = i < n_ent ? idx[first + i] : idx_base - 1;
    }
  };
  if (tid == 0) {
    mbar_init(s_full, 1);
    fence_mbar_init();
  }
  issue(tile);
  {
    uint4* z = reinterpret_cast<uint4*>(s_cell);
    for (int i = tid; i < n_cell16; i += kPartThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  unsigned parity = 0;
  for (; tile < n_tiles; tile += gridDim.x) {
    mbar_wait_parked(s_full, parity);
    parity ^= 1u;
    unsigned ent[kPartPer];
#pragma unroll
    for (int j = 0; j < kPartPer / 2; ++j) {
      const longlong2 t = s_raw[j * kPartThreads + tid];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long b = (h ? t.y : t.x) - idx_base;
        const unsigned lo = static_cast<unsigned>(b), hi = static_cast<unsigned>(static_cast<unsigned long long>(b) >> 32);
        const unsigned sel = (hi == 0u && lo < nb) ? lo : binless;
        unsigned short* c = s_cell + (sel >> kPartShift) * kPartThreads + cell;
        const unsigned old = *c;
        *c = static_cast<unsigned short>(old + 1u);
        ent[2 * j + h] = sel | (old << 22);
      }
    }
    __syncthreads();
    if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x);      // every thread has read its indices (barrier above)
    for (int b = warp; b <= n_buckets; b += kPartThreads / 32) {
      unsigned* c = reinterpret_cast<unsigned*>(s_cell + b * kPartThreads);
      unsigned w[kPartThreads / 64];
      unsigned run = 0;
#pragma unroll
      for (int j = 0; j < kPartThreads / 64; ++j) w[j] = c[j * 32 + lane];
#pragma unroll
      for (int j = 0; j < kPartThreads / 64; ++j) {
        const unsigned lo = w[j] & 0xffffu, hi = w[j] >> 16;
        w[j] = run | ((run + lo) << 16);
        run += lo + hi;
      }
      const unsigned incl = warp_inclusive_scan(run, lane);
      const unsigned base2 = (incl - run) * 0x10001u;
#pragma unroll
      for (int j = 0; j < kPartThreads / 64; ++j) c[j * 32 + lane] = w[j] + base2;
      if (lane == 31) s_tot[b] = static_cast<int>(incl);
    }
    __syncthreads();
    if (warp == 0) {
      const unsigned t = lane <= n_buckets ? static_cast<unsigned>(s_tot[lane]) : 0u;
      const unsigned incl = warp_inclusive_scan(t, lane);
      s_base[lane] = static_cast<int>(incl - t);
      if (lane == 31) s_base[32] = static_cast<int>(incl);
    }
    __syncthreads();
    if (tid <= n_buckets) offs[static_cast<long long>(tid) * n_tiles + tile] = static_cast<unsigned short>(s_base[tid]);
    if (tid < n_buckets && s_tot[tid] > 0)
      atomicAdd(totals + tid * kPartTotalStride, static_cast<unsigned long long>(s_tot[tid]));
#pragma unroll
    for (int j = 0; j < kPartPer; ++j) {
      const unsigned w = ent[j];
      const unsigned bucket = (w >> kPartShift) & 127u;
      const unsigned pos = static_cast<unsigned>(s_base[bucket]) + s_cell[bucket * kPartThreads + cell] + (w >> 22);
      s_sorted[pos] = static_cast<unsigned short>(w & (kPartWidth - 1));
    }
    __syncthreads();
    const int n16 = (s_base[n_buckets] * 2 + 15) >> 4;
    uint4* dst = reinterpret_cast<uint4*>(keys + tile * kPartTile);
    const uint4* src = reinterpret_cast<const uint4*>(s_sorted);
    for (int i = tid; i < n16; i += kPartThreads) __stcs(dst + i, src[i]);
    uint4* z = reinterpret_cast<uint4*>(s_cell);
    for (int i = tid; i < n_cell16; i += kPartThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kCountThreads, 1)
k_occurrence_bucket_kernel(const unsigned short* __restrict__ keys, const unsigned short* __restrict__ offs,
                           const unsigned long long* __restrict__ totals, int n_buckets, long long n_tiles,
                           int* __restrict__ counts, int lanes_env, int visit_cost) {
  extern __shared__ int s_bucket_hist[];      // kPartWidth counters
  __shared__ int s_first[kPartMaxBuckets + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    // the deal: a bucket with entries gets 1 + its share of the CTAs that are left over - the same on every CTA
    unsigned long long t[4], all = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = lane * 4 + i;
      t[i] = b < n_buckets ? totals[b * kPartTotalStride] : 0ull;
      if (t[i]) t[i] += static_cast<unsigned long long>(visit_cost) * static_cast<unsigned long long>(n_tiles);   // a visit costs like that many entries
      all += t[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) all += __shfl_xor_sync(kFull, all, o);
    const unsigned long long spare = gridDim.x - n_buckets;
    unsigned n[4], run = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      n[i] = t[i] ? 1u + static_cast<unsigned>(t[i] * spare / all) : 0u;
      run += n[i];
    }
    const unsigned incl = warp_inclusive_scan(run, lane);
    unsigned base = incl - run;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = lane * 4 + i;
      if (b < n_buckets) s_first[b] = static_cast<int>(base);
      base += n[i];
    }
    if (lane == 31) s_first[n_buckets] = static_cast<int>(incl);
  }
  for (int i = tid; i < kPartWidth; i += kCountThreads) s_bucket_hist[i] = 0;
  __syncthreads();
  const int me = blockIdx.x;
  if (me >= s_first[n_buckets]) return;
  int b = 0;
  {
    int lo = 0, hi = n_buckets;                 // the last bucket whose first CTA is <= me owns CTAs (next first > me)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_first[mid] <= me) lo = mid; else hi = mid;
    }
    b = lo;
  }
  const int s = me - s_first[b], n = s_first[b + 1] - s_first[b];
  const long long t_begin = n_tiles * s / n, t_end = n_tiles * (s + 1) / n;
  const unsigned short* o0 = offs + static_cast<long long>(b) * n_tiles;
  const unsigned short* o1 = o0 + n_tiles;
  const unsigned* keys32 = reinterpret_cast<const unsigned*>(keys);
  // warp w takes the tiles t_begin + w, + 32, + 64, ... (a hub-heavy bucket is cut into many CTAs of few tiles each:
  // every warp must get some); lane l fetches the segment bounds of the warp's l-th tile of a group of 32.  A warp
  // works on 32 / L segments at once, L lanes each: a thin bucket's segment (86 keys per tile for a 1 % bucket) would
  // leave most of a full warp's 128 word slots empty and cost one memory round trip per tile.
  constexpr int kWarps = kCountThreads / 32;
  const unsigned long long seg_words = totals[b * kPartTotalStride] / (2ull * static_cast<unsigned long long>(n_tiles));
  const int lanes_log2 = lanes_env > 0 ? lanes_env : (seg_words <= 40 ? 3 : seg_words <= 80 ? 4 : 5);
  const int sub = lane >> lanes_log2, sl = lane & ((1 << lanes_log2) - 1), nsub = 32 >> lanes_log2;
  const unsigned step = 4u << lanes_log2;
  for (long long g = t_begin + warp; g < t_end; g += 32ll * kWarps) {
    const long long t = g + static_cast<long long>(lane) * kWarps;
    unsigned st = 0, en = 0;
    if (t < t_end) {
      st = o0[t];
      en = o1[t];
    }
    const long long left = (t_end - g + kWarps - 1) / kWarps;
    const int nt = left < 32 ? static_cast<int>(left) : 32;
    for (int i = 0; i < nt; i += nsub) {                          // (warp-uniform trip count)
      const int mine = i + sub;                                   // <= 31; the lanes beyond the group's tiles hold 0, 0
      const unsigned a = __shfl_sync(kFull, st, mine), e = __shfl_sync(kFull, en, mine);
      const unsigned* kw = keys32 + (g + static_cast<long long>(mine) * kWarps) * (kPartTile / 2);
      const unsigned w_end = a < e ? (e + 1) >> 1 : 0u;
      for (unsigned w = (a >> 1) + sl; w < w_end; w += step) {     // (per-lane trip counts: no warp-wide operation inside)
        unsigned x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = w + (u << lanes_log2) < w_end ? __ldcs(kw + w + (u << lanes_log2)) : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned ww = w + (u << lanes_log2);
          if (ww < w_end) {
            // word ww holds keys 2 ww and 2 ww + 1; the first word may start one key early, the last end one late
            if (2 * ww >= a) atomicAdd(&s_bucket_hist[x[u] & 0xffffu], 1);
            if (2 * ww + 1 < e) atomicAdd(&s_bucket_hist[x[u] >> 16], 1);
          }
        }
      }
    }
  }
  __syncthreads();
  int* out = counts + (static_cast<long long>(b) << kPartShift);
  for (int i = tid; i < kPartWidth; i += kCountThreads) {
    const int c = s_bucket_hist[i];
    if (c) atomicAdd(out + i, c);
  }
}

// ------------------------------------------------------------------------------- retrieval metrics
// Recall@K / Precision@K / NDCG@K, reciprocal rank and average precision of every query from its
// ranked top-k list and its set of relevant items (src/utils/metrics.py:386-574: binary relevance,
// DCG = sum rel_i / log2(i + 2), IDCG with all relevant items first).  One warp per query: lane j
// tests ranks j and j + 32 against the relevant set, the hit mask is shared with two ballots.
__global__ void __launch_bounds__(128)
retrieval_metrics_kernel(const long long* __restrict__ topk, long long nq, int k,
                         const long long* __restrict__ rel_ptr, const long long* __restrict__ rel_idx,
                         const MetricKs ks, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long q = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const long long r0 = rel_ptr[q], r1 = rel_ptr[q + 1];
  const int n_rel = static_cast<int>(r1 - r0);
  unsigned hit[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    bool is_hit = false;
    if (j < k) {
      const long long id = topk[q * k + j];
      if (id >= 0)
        for (long long r = r0; r < r1 && !is_hit; ++r) is_hit = rel_idx[r] == id;
    }
    hit[h] = __ballot_sync(kFull, is_hit);
  }
  const unsigned long long mask = static_cast<unsigned long long>(hit[0]) | (static_cast<unsigned long long>(hit[1]) << 32);
  if (lane != 0) return;
  const int cols = 2 + 3 * ks.n;
  float* o = out + q * cols;
  double rr = 0.0, ap = 0.0;
  if (mask) rr = 1.0 / static_cast<double>(__ffsll(static_cast<long long>(mask)));
  {
    int seen = 0;
    unsigned long long m = mask;
    while (m) {
      const int p = __ffsll(static_cast<long long>(m)) - 1;
      m &= m - 1;
      ++seen;
      ap += static_cast<double>(seen) / static_cast<double>(p + 1);
    }
    ap = n_rel > 0 ? ap / n_rel : 0.0;
  }
  o[0] = static_cast<float>(rr);
  o[1] = static_cast<float>(ap);
  for (int t = 0; t < ks.n; ++t) {
    const int kv = ks.k[t];
    const unsigned long long lim = kv >= 64 ? ~0ull : ((1ull << kv) - 1ull);
    unsigned long long m = mask & lim;
    const int hits = __popcll(m);
    double dcg = 0.0, idcg = 0.0;
    while (m) {
      const int p = __ffsll(static_cast<long long>(m)) - 1;
      m &= m - 1;
      dcg += 1.0 / log2(static_cast<double>(p + 2));
    }
    const int ideal = n_rel < kv ? n_rel : kv;
    for (int i = 0; i < ideal; ++i) idcg += 1.0 / log2(static_cast<double>(i + 2));
    o[2 + t] = n_rel > 0 ? static_cast<float>(static_cast<double>(hits) / n_rel) : 0.f;
    o[2 + ks.n + t] = kv > 0 ? static_cast<float>(static_cast<double>(hits) / kv) : 0.f;
    o[2 + 2 * ks.n + t] = idcg > 0.0 ? static_cast<float>(dcg / idcg) : 0.f;
  }
}

// ------------------------------------------------------------------------------- gather rows
__global__ void gather_rows_kernel(const float* __restrict__ g_f32,
                                   const __nv_bfloat16* __restrict__ g_bf16, int d, int d_pad,
                                   const long long* __restrict__ idx, long long n, long long n_rows,
                                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const long long gi = idx[r];
  for (int i = lane; i < d; i += 32) {
    float x = 0.f;
    if (gi >= 0 && gi < n_rows)
      x = g_f32 ? g_f32[gi * d + i] : __bfloat162float(g_bf16[gi * d_pad + i]);
    out[r * d + i] = x;
  }
}

}  // namespace

// =============================================================================== launchers
cudaError_t launch_prep_rows_bcast(const void* rows, int dtype, int64_t n, int d, int d_pad, bool normalize,
                                   const BcastSpec& dst, int64_t dst_row0, float* out_f32, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  long long blocks = (n * 32 + block - 1) / block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int grid = static_cast<int>(blocks);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  bool vec = dtype == TVC_F32 && !normalize && (d % 8) == 0 && al16(rows) && (!out_f32 || al16(out_f32));
  for (int t = 0; t < dst.n; ++t) vec = vec && al16(dst.bf16[t]);
  if (vec) {
    prep_rows_f32v_kernel<<<grid, block, 0, stream>>>(static_cast<const float*>(rows), n, d, d_pad, dst, dst_row0,
                                                      out_f32);
    note_launch();
    return cudaGetLastError();
  }
  // general path: one launch per destination (rows are re-read; only odd shapes / dtypes get here)
  for (int t = 0; t < (dst.n > 0 ? dst.n : 1); ++t) {
    __nv_bfloat16* ob = dst.n > 0 ? dst.bf16[t] + dst_row0 * d_pad : nullptr;
    float* of = t == 0 ? out_f32 : nullptr;
    switch (dtype) {
      case TVC_F32:
        prep_rows_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(rows), n, d, d_pad, normalize,
                                                            ob, of);
        break;
      case TVC_BF16:
        prep_rows_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(rows), n, d,
                                                                    d_pad, normalize, ob, of);
        break;
      case TVC_F16:
        prep_rows_kernel<__half><<<grid, block, 0, stream>>>(static_cast<const __half*>(rows), n, d, d_pad,
                                                             normalize, ob, of);
        break;
      default:
        return cudaErrorInvalidValue;
    }
    note_launch();
  }
  return cudaGetLastError();
}

cudaError_t launch_prep_rows(const void* rows, int dtype, int64_t n, int d, int d_pad, bool normalize,
                             __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t stream) {
  BcastSpec dst{};
  if (out_bf16) {
    dst.n = 1;
    dst.bf16[0] = out_bf16;
  }
  return launch_prep_rows_bcast(rows, dtype, n, d, d_pad, normalize, dst, 0, out_f32, stream);
}

cudaError_t launch_rerank(const float* cand_val, const int32_t* cand_idx, int64_t m, int64_t full_rows, int splits,
                          int kp, int k, const float* q_f32, const float* g_f32, int d,
                          float threshold, int64_t global_row_offset, float* out_sim,
                          int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  rerank_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(
      cand_val, cand_idx, m, full_rows, splits, kp, k, q_f32, g_f32, d, threshold, global_row_offset, out_sim,
      reinterpret_cast<long long*>(out_idx));
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_select_candidates(const float* cand_val, const int32_t* cand_idx, int64_t m, int64_t full_rows,
                                     int splits, int kp, int64_t global_row_offset, const ScatterSpec& sc, float* out_val,
                                     int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  select_candidates_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(cand_val, cand_idx, m, full_rows, splits, kp,
                                                                   global_row_offset, sc, out_val,
                                                                   reinterpret_cast<long long*>(out_idx));
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_rerank_merged(const float* cand_val, const int64_t* cand_idx, int64_t m, int parts, int kp,
                                 int k, const float* q_f32, const RowSource& src, int d, float threshold,
                                 float* out_sim, int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  rerank_merged_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(
      cand_val, reinterpret_cast<const long long*>(cand_idx), m, parts, kp, k, q_f32, src, d, threshold, out_sim,
      reinterpret_cast<long long*>(out_idx));
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_exchange_merge(const float* cand_val, const int64_t* cand_idx, int64_t m, int parts, int kp,
                                  const ReqDst& dst, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  exchange_merge_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(cand_val, reinterpret_cast<const long long*>(cand_idx),
                                                               m, parts, kp, dst);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_exchange_rescore(const int64_t* req, const float* q_f32, const float* g_f32, int64_t g_off,
                                    int64_t g_n, int d, int owners, int64_t rows_per_slice, int64_t m_total, int kp,
                                    const ScoreDst& dst, cudaStream_t stream) {
  if (m_total <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m_total + kRerankWarps - 1) / kRerankWarps);
  exchange_rescore_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(reinterpret_cast<const long long*>(req), q_f32, g_f32,
                                                                 g_off, g_n, d, owners, rows_per_slice, m_total, kp, dst);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_exchange_finalize(const int64_t* req, const float* score, int64_t m, int kp, int k, float threshold,
                                     float* out_sim, int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  exchange_finalize_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(reinterpret_cast<const long long*>(req), score, m, kp,
                                                                  k, threshold, out_sim,
                                                                  reinterpret_cast<long long*>(out_idx));
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_merge_topk(const float* in_sim, const int64_t* in_idx, int64_t m, int parts, int k,
                              float* out_sim, int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  const int grid = static_cast<int>((m + 3) / 4);
  merge_topk_kernel<<<grid, 128, 0, stream>>>(in_sim, reinterpret_cast<const long long*>(in_idx), m,
                                              parts, k, out_sim,
                                              reinterpret_cast<long long*>(out_idx));
  note_launch();
  return cudaGetLastError();
}

namespace {
struct PartLayout {
  long long n_tiles;
  int n_buckets;
  size_t totals_b, offs_b, keys_b;
};
PartLayout part_layout(long long total, long long n_bins) {
  PartLayout l;
  l.n_tiles = (total + kPartTile - 1) / kPartTile;
  l.n_buckets = static_cast<int>((n_bins + kPartWidth - 1) >> kPartShift);
  auto up = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  l.totals_b = up(static_cast<size_t>(l.n_buckets) * kPartTotalStride * 8);
  l.offs_b = up(static_cast<size_t>(l.n_buckets + 1) * l.n_tiles * 2);
  l.keys_b = up(static_cast<size_t>(l.n_tiles) * kPartTile * 2);
  return l;
}
}  // namespace

size_t k_occurrence_part_scratch_bytes(const int64_t* idx, int64_t m, int k, int64_t n_bins, int64_t part_min) {
  const long long total = m * k;
  if (total <= 0 || total < part_min) return 0;
  if (n_bins * 4 <= 48 * 1024 && total >= 64 * n_bins) return 0;                  // the shared-memory path
  if (n_bins > static_cast<long long>(kPartMaxBuckets) << kPartShift) return 0;
  if ((reinterpret_cast<uintptr_t>(idx) & 15u) != 0) return 0;
  if ((total + kPartTile - 1) / kPartTile > 0x7fffffffll) return 0;
  const PartLayout l = part_layout(total, n_bins);
  return l.totals_b + l.offs_b + l.keys_b;
}

cudaError_t launch_k_occurrence(const int64_t* idx, int64_t m, int k, int64_t idx_base,
                                int64_t n_bins, int32_t* counts, int sm_count, int* hot_scratch,
                                void* part_scratch, cudaStream_t stream) {
  const long long total = m * k;
  if (total <= 0 || n_bins <= 0) return cudaSuccess;
  if (part_scratch != nullptr) {
    // bucketed path (the caller sized part_scratch with k_occurrence_part_scratch_bytes, 256-byte aligned)
    const PartLayout l = part_layout(total, n_bins);
    uint8_t* base = static_cast<uint8_t*>(part_scratch);
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(base);
    unsigned short* offs = reinterpret_cast<unsigned short*>(base + l.totals_b);
    unsigned short* keys = reinterpret_cast<unsigned short*>(base + l.totals_b + l.offs_b);
    cudaError_t e = cudaMemsetAsync(totals, 0, l.totals_b, stream);
    if (e != cudaSuccess) return e;
    static SmemAttrOnce attr_part2, attr_part3, attr_count;
    constexpr int kMaxSmem = 227 * 1024 - 2048;      // dynamic part: the kernels hold up to 1 KB of static tables
    if ((e = attr_part2.ensure(reinterpret_cast<const void*>(k_occurrence_partition_kernel<2>), kMaxSmem)) != cudaSuccess) return e;
    if ((e = attr_part3.ensure(reinterpret_cast<const void*>(k_occurrence_partition_kernel<3>), kMaxSmem)) != cudaSuccess) return e;
    if ((e = attr_count.ensure(reinterpret_cast<const void*>(k_occurrence_bucket_kernel), kMaxSmem)) != cudaSuccess) return e;
    const size_t smem1 = static_cast<size_t>(l.n_buckets + 1) * kPartThreads * 2 + static_cast<size_t>(kPartTile) * 2;
    // measurements: TVC_KOCC_PART_OCC=2|3 pins the register budget (CTAs per SM), TVC_KOCC_PART_KIND=1 the general
    // (shared-memory cell) partition kernel where the vote kernel would run
    static const int occ_env = [] {
      const char* v = getenv("TVC_KOCC_PART_OCC");
      return v ? atoi(v) : 0;
    }();
    static const int kind_env = [] {
      const char* v = getenv("TVC_KOCC_PART_KIND");
      return v ? atoi(v) : 0;
    }();
    const long long* ip = reinterpret_cast<const long long*>(idx);
    long long tile0 = 0;
    if (l.n_buckets < kTmaCellBuckets && kind_env != 1) {
      // persistent, copy-engine fed
      static SmemAttrOnce attr_tma;
      if ((e = attr_tma.ensure(reinterpret_cast<const void*>(k_occurrence_partition_tma_kernel),
                               static_cast<int>(kTmaSmemBytes))) != cudaSuccess) return e;
      const long long cap = 2ll * sm_count;
      const unsigned g = static_cast<unsigned>(l.n_tiles < cap ? l.n_tiles : cap);
      k_occurrence_partition_tma_kernel<<<g, kPartThreads, kTmaSmemBytes, stream>>>(ip, total, idx_base, n_bins, l.n_buckets,
                                                                                  l.n_tiles, keys, offs, totals);
      note_launch();
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      tile0 = l.n_tiles;
    }
    if (tile0 < l.n_tiles) {
      const unsigned g = static_cast<unsigned>(l.n_tiles - tile0);
      const bool three = occ_env ? occ_env >= 3 : (smem1 + 2048) * 3 <= static_cast<size_t>(227 * 1024);
      if (three)
        k_occurrence_partition_kernel<3><<<g, kPartThreads, smem1, stream>>>(ip, total, idx_base, n_bins, l.n_buckets,
                                                                            l.n_tiles, keys, offs, totals, tile0);
      else
        k_occurrence_partition_kernel<2><<<g, kPartThreads, smem1, stream>>>(ip, total, idx_base, n_bins, l.n_buckets,
                                                                            l.n_tiles, keys, offs, totals, tile0);
      note_launch();
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // CTAs of pass 2: each one zeroes and flushes 32768 counters, so few and long ones (TVC_KOCC_PART_GRID = CTAs per SM)
    static const int grid_env = [] {
      const char* v = getenv("TVC_KOCC_PART_GRID");
      return v ? atoi(v) : 0;
    }();
    int grid2 = sm_count * (grid_env > 0 ? grid_env : 3);      // measured 1 / 2 / 3 / 4 per SM: 251 / 228 / 222 / 231 us
    if (grid2 < 2 * l.n_buckets) grid2 = 2 * l.n_buckets;
    // measurements: TVC_KOCC_PART_LANES = log2 of the lanes per segment (3..5; default by segment length),
    // TVC_KOCC_PART_VISIT = what a (bucket, tile) visit weighs in the deal, in entries
    static const int lanes_env = [] {
      const char* v = getenv("TVC_KOCC_PART_LANES");
      const int x = v ? atoi(v) : 0;
      return x >= 3 && x <= 5 ? x : 0;
    }();
    static const int visit_env = [] {
      const char* v = getenv("TVC_KOCC_PART_VISIT");
      return v ? atoi(v) : -1;
    }();
    k_occurrence_bucket_kernel<<<grid2, kCountThreads, kPartWidth * 4, stream>>>(keys, offs, totals, l.n_buckets,
                                                                                 l.n_tiles, counts, lanes_env,
                                                                                 visit_env >= 0 ? visit_env : 128);   // measured 0 / 128 / 512: 204.7 / 200.0 / 202.9 us
    note_launch();
    return cudaGetLastError();
  }
  const int block = 256;
  const long long quads = (total + 3) / 4;
  long long blocks = (quads + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count) * 5;   // 38 KB of tables per block
  if (blocks > cap) blocks = cap;
  const size_t hist_bytes = static_cast<size_t>(n_bins) * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(idx) & 15u) == 0;
  const long long* ip = reinterpret_cast<const long long*>(idx);
  // private shared-memory histograms pay off when a block sees many increments per bin it flushes
  const bool use_smem = hist_bytes <= 48 * 1024 && total >= 64 * n_bins;
  if (use_smem) {
    long long b2 = total / (16 * n_bins);       // every block flushes up to n_bins counters
    if (b2 < blocks) blocks = b2 < 1 ? 1 : b2;
    const int g = static_cast<int>(blocks);
    if (aligned)
      k_occurrence_smem_kernel<true><<<g, block, hist_bytes, stream>>>(ip, total, idx_base, n_bins, counts);
    else
      k_occurrence_smem_kernel<false><<<g, block, hist_bytes, stream>>>(ip, total, idx_base, n_bins, counts);
    note_launch();
    return cudaGetLastError();
  }
  // hot-bin list: a key table in per-stream scratch written by the sampling pass, read by the main kernel
  int* hot = nullptr;
  if (hot_scratch != nullptr && total >= 4096 && n_bins <= (1ll << 29)) {   // line numbers fit 24 bits
    hot = hot_scratch;
    k_occurrence_sample_kernel<<<1, 1024, 0, stream>>>(ip, total, idx_base, n_bins, hot);
    note_launch();
  }
  // TVC_KOCC_AGG=0/1 pins the warp vote off / on (measurements); default: the pre-pass decides
  static const int force_agg = [] {
    const char* e = getenv("TVC_KOCC_AGG");
    return e ? atoi(e) : -1;
  }();
  const int g = static_cast<int>(blocks);
  if (aligned)
    k_occurrence_kernel<true><<<g, block, 0, stream>>>(ip, total, idx_base, n_bins, counts, hot, force_agg);
  else
    k_occurrence_kernel<false><<<g, block, 0, stream>>>(ip, total, idx_base, n_bins, counts, hot, force_agg);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_retrieval_metrics(const int64_t* topk, int64_t nq, int k, const int64_t* rel_ptr,
                                     const int64_t* rel_idx, const MetricKs& ks, float* out, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int grid = static_cast<int>((nq + 3) / 4);
  retrieval_metrics_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const long long*>(topk), nq, k,
                                                     reinterpret_cast<const long long*>(rel_ptr),
                                                     reinterpret_cast<const long long*>(rel_idx), ks, out);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* g_f32, const __nv_bfloat16* g_bf16, int d, int d_pad,
                               const int64_t* idx, int64_t n, int64_t n_rows, float* out,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  const int grid = static_cast<int>((n * 32 + block - 1) / block);
  gather_rows_kernel<<<grid, block, 0, stream>>>(g_f32, g_bf16, d, d_pad,
                                                 reinterpret_cast<const long long*>(idx), n, n_rows,
                                                 out);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tvc
