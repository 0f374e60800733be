// Bandwidth-bound kernels around the GEMM: operand preparation, candidate re-rank / merge,
// kernel (b) variant-consistency reduction and kernel (c) k-occurrence histogram.  128-bit loads,
// warp shuffles, shared-memory staging, warp-aggregated atomics; no tensor cores.
#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>

#include <atomic>

#include "tvc_internal.h"

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
              int splits, int kp, int k, const float* __restrict__ q_f32,
              const float* __restrict__ g_f32, int d, float threshold, long long row_offset,
              float* __restrict__ out_sim, long long* __restrict__ out_idx) {
  __shared__ float s_val[kRerankWarps][kMaxKp];
  __shared__ int s_idx[kRerankWarps][kMaxKp];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRerankWarps + w;
  if (row >= m) return;
  const int ncand = splits * kp;
  const float* cv = cand_val + row * ncand;
  const int32_t* ci = cand_idx + row * ncand;

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
// The int64 index stream is read with 128-bit loads at HBM speed (5.9 TB/s measured); what bounds the
// kernel is the atomic path.  Measured on B200 (50M entries, 1M bins): one RED per entry sustains
// 180 G entries/s when bins are spread, but a single hot bin serialises at ~1.4 G/s in L2 - and
// hubness histograms are exactly the case with hot bins (an adversarial hub is in most rows).  So:
//   * a sampling pre-pass (one block, 8192 strided entries, shared-memory hash counts) decides whether
//     any bin holds >= 1/128 of the stream;
//   * spread data  -> plain RED per entry, four entries per thread in flight;
//   * hot bins     -> warp-aggregated atomics: lanes holding the same bin (__match_any_sync) elect a
//     leader that issues one atomic of the group size (4.6x faster on one-hub-per-row data, 2.6x
//     slower on spread data, hence the switch);
//   * histograms that fit in shared memory (<= 48 KB) accumulate privately per block and flush once.
__global__ void __launch_bounds__(1024)
k_occurrence_sample_kernel(const long long* __restrict__ idx, long long total, int* __restrict__ hot_flag) {
  __shared__ int s_cnt[4096];
  __shared__ int s_max;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_cnt[i] = 0;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const int samples = total < 8192 ? static_cast<int>(total) : 8192;
  const long long stride = total / samples;
  for (int i = threadIdx.x; i < samples; i += blockDim.x) {
    const long long b = idx[static_cast<long long>(i) * stride];
    if (b >= 0) {
      // mix the bits so that neighbouring bins do not share a slot with a hot one
      const unsigned h = static_cast<unsigned>((static_cast<unsigned long long>(b) * 0x9E3779B97F4A7C15ull) >> 52);
      atomicAdd(&s_cnt[h], 1);
    }
  }
  __syncthreads();
  int m = 0;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) m = max(m, s_cnt[i]);
  atomicMax(&s_max, m);
  __syncthreads();
  if (threadIdx.x == 0) *hot_flag = (s_max * 128 >= samples && s_max >= 4) ? 1 : 0;
}

template <bool SMEM, bool VEC>
__global__ void __launch_bounds__(256)
k_occurrence_kernel(const long long* __restrict__ idx, long long total, long long idx_base,
                    long long n_bins, int* __restrict__ counts, const int* __restrict__ hot_flag) {
  extern __shared__ int s_hist[];
  if (SMEM) {
    for (long long b = threadIdx.x; b < n_bins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
  }
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool aggregate = SMEM || (hot_flag != nullptr && *hot_flag != 0);   // block-uniform
  if (!aggregate) {
    // spread bins: one RED per entry, four entries (two 128-bit loads) per thread per trip
    const long long quads = VEC ? (total >> 2) : 0;
    const longlong2* idx2 = reinterpret_cast<const longlong2*>(idx);
    for (long long p = tid; p < quads; p += nthreads) {
      const longlong2 a = idx2[2 * p], c = idx2[2 * p + 1];
      const long long b[4] = {a.x - idx_base, a.y - idx_base, c.x - idx_base, c.y - idx_base};
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if (b[h] >= 0 && b[h] < n_bins) atomicAdd(&counts[b[h]], 1);
    }
    for (long long e = (quads << 2) + tid; e < total; e += nthreads) {
      const long long b = idx[e] - idx_base;
      if (b >= 0 && b < n_bins) atomicAdd(&counts[b], 1);
    }
    return;
  }
  const long long pairs = total >> 1;
  const longlong2* idx2 = reinterpret_cast<const longlong2*>(idx);
  // every lane of a warp takes the same number of trips (the ballot below names the participants)
  for (long long p0 = tid - (threadIdx.x & 31); p0 < pairs + 1; p0 += nthreads) {
    const long long p = p0 + (threadIdx.x & 31);
    long long b0 = -1, b1 = -1;
    if (p < pairs) {
      longlong2 v;
      if (VEC) {
        v = idx2[p];
      } else {
        v.x = idx[2 * p];
        v.y = idx[2 * p + 1];
      }
      b0 = v.x - idx_base;
      b1 = v.y - idx_base;
    } else if (p == pairs && (total & 1)) {
      b0 = idx[total - 1] - idx_base;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long b = h == 0 ? b0 : b1;
      const bool valid = b >= 0 && b < n_bins;
      const unsigned act = __ballot_sync(kFull, valid);
      if (valid) {
        const unsigned peers = __match_any_sync(act, static_cast<unsigned long long>(b));
        if ((threadIdx.x & 31) == __ffs(peers) - 1) {
          if (SMEM)
            atomicAdd(&s_hist[b], __popc(peers));
          else
            atomicAdd(&counts[b], __popc(peers));
        }
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    for (long long b = threadIdx.x; b < n_bins; b += blockDim.x) {
      const int c = s_hist[b];
      if (c) atomicAdd(&counts[b], c);
    }
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

// ------------------------------------------------------------------------------- kernel (b)
// Statistics and decisions in fp64 on fp32 similarities, in the reference's operation order
// (Python floats -> np.mean / np.std(ddof=0) / np.var).
struct Stats {
  double mean, var, sd, mn, mx;
};
__device__ __forceinline__ Stats stats_of(const float* x, int n) {
  Stats s{0., 0., 0., 0., 0.};
  if (n <= 0) return s;
  double sum = 0., mn = x[0], mx = x[0];
  for (int i = 0; i < n; ++i) {
    const double v = x[i];
    sum += v;
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  s.mean = sum / n;
  double acc = 0.;
  for (int i = 0; i < n; ++i) {
    const double dlt = static_cast<double>(x[i]) - s.mean;
    acc += dlt * dlt;
  }
  s.var = acc / n;
  s.sd = sqrt(s.var);
  s.mn = mn;
  s.mx = mx;
  return s;
}

__device__ __forceinline__ double clipd(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

// sv/sr/sg/sx point at this query's similarity lists (any address space).
__device__ void finish_scores(const tvc_detector_params& p, float s0f, const float* sv, int nv,
                              const float* sr, int nr, const float* sg, int ng, const float* sx,
                              int nx, float* out /*[TVC_NSCORES]*/, uint8_t* flag) {
  const double s0 = s0f;
  const Stats tv = stats_of(sv, nv);
  const Stats rt = stats_of(sr, nr);
  const Stats gn = stats_of(sg, ng);
  const Stats xv = stats_of(sx, nx);

  // --- AdversarialDetector (src/detector.py:441-590, 643-682, 399)
  double det_tv = 0.0;
  if (nv > 0) {
    const double consistency = 1.0 - fabs(s0 - tv.mean);
    const double variability = 1.0 - tv.sd;
    det_tv = 1.0 - (consistency * 0.7 + variability * 0.3);
  }
  const double det_sd = ng > 0 ? 1.0 - gn.mean : 0.0;
  const double det_c = 1.0 - s0;
  double agg = 0.0;
  {
    const double sc[3] = {det_tv, det_sd, det_c};
    const double wt[3] = {p.w_text_variants, p.w_sd_reference, p.w_consistency};
    double wsum = 0., tw = 0., sum = 0., mx = -INFINITY, mn = INFINITY;
    int cnt = 0;
    for (int i = 0; i < 3; ++i) {
      if (!(p.methods & (1u << i))) continue;
      wsum += sc[i] * wt[i];
      tw += wt[i];
      sum += sc[i];
      mx = sc[i] > mx ? sc[i] : mx;
      mn = sc[i] < mn ? sc[i] : mn;
      ++cnt;
    }
    if (cnt > 0) {
      if (p.aggregation == 0)
        agg = tw > 0. ? wsum / tw : 0.0;
      else if (p.aggregation == 2)
        agg = mx;
      else if (p.aggregation == 3)
        agg = mn;
      else
        agg = sum / cnt;
    }
  }
  const bool det_adv = agg > static_cast<double>(p.detection_threshold);

  // --- MultiModalDefenseDetector scores (experiments/defenses/detector.py:228-300)
  const double tv_c = nv > 0 ? tv.mean : s0;
  const double tv_s = nv > 0 ? tv.sd : 0.0;
  const double rt_c = nr > 0 ? rt.mean : 0.0, rt_s = nr > 0 ? rt.sd : 0.0;
  const double gn_c = ng > 0 ? gn.mean : 0.0, gn_s = ng > 0 ? gn.sd : 0.0;
  const double four[4] = {s0, tv_c, rt_c, gn_c};
  double valid[4];
  int nvalid = 0;
  for (int i = 0; i < 4; ++i)
    if (four[i] > 0.) valid[nvalid++] = four[i];
  double vmean = 0., vvar = 0.;
  if (nvalid > 0) {
    for (int i = 0; i < nvalid; ++i) vmean += valid[i];
    vmean /= nvalid;
    for (int i = 0; i < nvalid; ++i) vvar += (valid[i] - vmean) * (valid[i] - vmean);
    vvar /= nvalid;
  }
  const double cmv = nvalid < 2 ? 0.0 : vvar;

  // --- ConsistencyChecker (experiments/defenses/consistency_checker.py:119-272)
  double overall = 0.0;
  if (p.voting == 0) {
    overall = nvalid > 0 ? vmean : 0.0;
  } else {
    double w[4];
    if (p.voting == 1) {
      for (int i = 0; i < 4; ++i) w[i] = p.cc_weights[i];
    } else {
      w[0] = 1.0;
      w[1] = 1.0 / (1.0 + tv_s);
      w[2] = 1.0 / (1.0 + rt_s);
      w[3] = 1.0 / (1.0 + gn_s);
      const double t = w[0] + w[1] + w[2] + w[3];
      if (t > 0.)
        for (int i = 0; i < 4; ++i) w[i] /= t;
    }
    double ws = 0., tw = 0.;
    for (int i = 0; i < 4; ++i)
      if (four[i] > 0.) {
        ws += four[i] * w[i];
        tw += w[i];
      }
    overall = tw == 0. ? 0.0 : ws / tw;
  }
  double thr = p.cc_base_threshold;
  if (p.cc_adaptive) {
    if (cmv > 0.1) thr += 0.1;
    const double avg_std = (tv_s + rt_s + gn_s) / 3.0;
    if (avg_std > 0.2) thr += 0.05;
    thr = clipd(thr, 0.1, 0.9);
  }
  const bool cc_adv = overall < thr;
  const double dist_conf = fabs(overall - thr) / thr;
  const double cons_conf = nvalid > 1 ? 1.0 - sqrt(vvar) : 0.5;
  const double var_conf = 1.0 - (cmv < 1.0 ? cmv : 1.0);
  const double conf = clipd((dist_conf + cons_conf + var_conf) / 3.0, 0.0, 1.0);

  // --- README sigma rule over all references (README.md:474-482, 846)
  double sigma = 0.0;
  {
    const int n = nr + ng;
    if (n > 0) {
      double sum = 0.;
      for (int i = 0; i < nr; ++i) sum += sr[i];
      for (int i = 0; i < ng; ++i) sum += sg[i];
      const double mu = sum / n;
      double acc = 0.;
      for (int i = 0; i < nr; ++i) acc += (sr[i] - mu) * (sr[i] - mu);
      for (int i = 0; i < ng; ++i) acc += (sg[i] - mu) * (sg[i] - mu);
      sigma = sqrt(acc / n);
    }
  }
  const bool sig_adv = sigma > static_cast<double>(p.sigma_threshold);

  out[TVC_S_ORIGINAL] = s0f;
  out[TVC_S_TV_MEAN] = static_cast<float>(tv_c);
  out[TVC_S_TV_STD] = static_cast<float>(tv_s);
  out[TVC_S_TV_MIN] = static_cast<float>(nv > 0 ? tv.mn : s0);
  out[TVC_S_TV_VAR] = static_cast<float>(nv > 0 ? tv.var : 0.0);
  out[TVC_S_RET_MEAN] = static_cast<float>(rt_c);
  out[TVC_S_RET_STD] = static_cast<float>(rt_s);
  out[TVC_S_GEN_MEAN] = static_cast<float>(gn_c);
  out[TVC_S_GEN_STD] = static_cast<float>(gn_s);
  out[TVC_S_GEN_MAX] = static_cast<float>(ng > 0 ? gn.mx : 0.0);
  out[TVC_S_CROSS_MODAL_VAR] = static_cast<float>(cmv);
  out[TVC_S_XV_MEAN] = static_cast<float>(xv.mean);
  out[TVC_S_XV_MIN] = static_cast<float>(xv.mn);
  out[TVC_S_XV_VAR] = static_cast<float>(xv.var);
  out[TVC_S_DET_TV] = static_cast<float>(det_tv);
  out[TVC_S_DET_SD] = static_cast<float>(det_sd);
  out[TVC_S_DET_C] = static_cast<float>(det_c);
  out[TVC_S_DET_AGG] = static_cast<float>(agg);
  out[TVC_S_CC_OVERALL] = static_cast<float>(overall);
  out[TVC_S_CC_THRESHOLD] = static_cast<float>(thr);
  out[TVC_S_CC_CONFIDENCE] = static_cast<float>(conf);
  out[TVC_S_N_RET] = static_cast<float>(nr);
  out[TVC_S_N_GEN] = static_cast<float>(ng);
  out[TVC_S_REF_SIGMA] = static_cast<float>(sigma);
  *flag = static_cast<uint8_t>((det_adv ? TVC_FLAG_DET_ADV : 0u) | (cc_adv ? TVC_FLAG_CC_ADV : 0u) |
                               (sig_adv ? TVC_FLAG_SIGMA_ADV : 0u));
}

// Similarity-fed mode: a block stages the contiguous similarity slabs of its 128 queries into
// shared memory with coalesced 128-bit loads, one thread reduces one query, results leave through
// shared memory as coalesced stores.
constexpr int kSimsBlock = 128;

__device__ __forceinline__ void stage_slab(float* dst, const float* __restrict__ src, long long q0,
                                           int nq, int width) {
  if (src == nullptr || width == 0) return;
  const long long base = q0 * width;
  const int total = nq * width;
  const float* s = src + base;
  if (((reinterpret_cast<uintptr_t>(s) & 15u) == 0)) {
    const int n4 = total >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) dst[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = s[i];
  }
}

__global__ void __launch_bounds__(kSimsBlock)
consistency_sims_kernel(const tvc_detector_params p, long long nq_total, const float* __restrict__ s0,
                        const float* __restrict__ sv, const float* __restrict__ sr,
                        const int32_t* __restrict__ r_cnt, const float* __restrict__ sg,
                        const int32_t* __restrict__ g_cnt, const float* __restrict__ sxv,
                        float* __restrict__ scores, uint8_t* __restrict__ flags) {
  extern __shared__ __align__(16) float s_buf[];
  const int V = p.n_variants, R = p.n_retrieval, G = p.n_generative;
  const int X = sxv ? V * (V - 1) / 2 : 0;
  const int pad4 = 4;  // keep every slab 16-byte aligned
  auto up4 = [](int x) { return (x + 3) & ~3; };
  float* b_sv = s_buf;
  float* b_sr = b_sv + up4(kSimsBlock * V) + pad4;
  float* b_sg = b_sr + up4(kSimsBlock * R) + pad4;
  float* b_sx = b_sg + up4(kSimsBlock * G) + pad4;
  float* b_out = b_sx + up4(kSimsBlock * X) + pad4;
  const long long q0 = static_cast<long long>(blockIdx.x) * kSimsBlock;
  const int nq = static_cast<int>(min(static_cast<long long>(kSimsBlock), nq_total - q0));
  stage_slab(b_sv, sv, q0, nq, V);
  stage_slab(b_sr, sr, q0, nq, R);
  stage_slab(b_sg, sg, q0, nq, G);
  stage_slab(b_sx, sxv, q0, nq, X);
  __syncthreads();
  const int t = threadIdx.x;
  uint8_t flag = 0;
  if (t < nq) {
    const long long q = q0 + t;
    const int nr = sr ? (r_cnt ? max(0, min(R, r_cnt[q])) : R) : 0;
    const int ng = sg ? (g_cnt ? max(0, min(G, g_cnt[q])) : G) : 0;
    const int nv = sv ? V : 0;
    float o[TVC_NSCORES];
    finish_scores(p, s0[q], b_sv + t * V, nv, b_sr + t * R, nr, b_sg + t * G, ng, b_sx + t * X, X, o,
                  &flag);
#pragma unroll
    for (int j = 0; j < TVC_NSCORES; ++j) b_out[t * TVC_NSCORES + j] = o[j];
    flags[q] = flag;
  }
  __syncthreads();
  float* dst = scores + q0 * TVC_NSCORES;
  const int total = nq * TVC_NSCORES;  // multiple of 4 and 16-byte aligned (24 floats per query)
  for (int i = threadIdx.x; i < (total >> 2); i += blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(b_out)[i];
}

// Embedding-fed mode: one warp per query.  Rows are read with 128-bit loads; the query image row
// and the kept reference rows live in shared memory for the de-duplication dots.
struct CosAcc {
  float dot, na, nb;
};
__device__ __forceinline__ CosAcc warp_cos_acc(const float* __restrict__ a,
                                               const float* __restrict__ b, int d) {
  const int lane = threadIdx.x & 31;
  float dot = 0.f, na = 0.f, nb = 0.f;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int i = lane; i < (d >> 2); i += 32) {
      const float4 x = a4[i], y = b4[i];
      dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot);
      dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
      na = fmaf(x.x, x.x, na); na = fmaf(x.y, x.y, na);
      na = fmaf(x.z, x.z, na); na = fmaf(x.w, x.w, na);
      nb = fmaf(y.x, y.x, nb); nb = fmaf(y.y, y.y, nb);
      nb = fmaf(y.z, y.z, nb); nb = fmaf(y.w, y.w, nb);
    }
  } else {
    for (int i = lane; i < d; i += 32) {
      const float x = a[i], y = b[i];
      dot = fmaf(x, y, dot);
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
  }
  CosAcc r;
  r.dot = warp_sum(dot);
  r.na = warp_sum(na);
  r.nb = warp_sum(nb);
  return r;
}
// torch.cosine_similarity: x.y / max(|x||y|, eps), eps = 1e-8
__device__ __forceinline__ float cos_from(const CosAcc& c) {
  const float den = fmaxf(sqrtf(c.na) * sqrtf(c.nb), 1e-8f);
  return c.dot / den;
}

__device__ __forceinline__ void load_row_to_smem(float* dst, const float* g_f32,
                                                 const __nv_bfloat16* g_bf16, int d, int d_pad,
                                                 long long gi) {
  const int lane = threadIdx.x & 31;
  if (g_f32) {
    const float* src = g_f32 + gi * d;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      for (int i = lane; i < (d >> 2); i += 32)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    } else {
      for (int i = lane; i < d; i += 32) dst[i] = src[i];
    }
  } else {
    const __nv_bfloat16* src = g_bf16 + gi * d_pad;
    for (int i = lane; i < d; i += 32) dst[i] = __bfloat162float(src[i]);
  }
  __syncwarp();
}

// Greedy reference selection (experiments/defenses/detector.py:184-204, 302-325): walk the
// candidate list in order, drop repeated indices and rows whose cosine to an already kept row
// exceeds dedup_threshold, stop at `cap` kept rows; sims[j] = cos(image, kept row j).
__device__ int select_refs(const float* s_img, float* s_rows, int d, const RowSource& src,
                           const long long* cand, int ncand, int cap, float dedup_thr, float* sims,
                           long long* kept_idx) {
  int kept = 0;
  for (int c = 0; c < ncand && kept < cap; ++c) {
    const long long gi = cand[c];
    int part = -1;
    for (int p = 0; p < src.nparts; ++p)
      if (gi >= src.off[p] && gi < src.off[p] + src.n[p]) part = p;
    if (part < 0) continue;  // unused slot (-1) or an index no shard owns
    bool dup = false;
    for (int j = 0; j < kept; ++j) dup |= (kept_idx[j] == gi);
    if (dup) continue;
    float* row = s_rows + static_cast<size_t>(kept) * d;
    load_row_to_smem(row, src.f32[part], src.bf16[part], d, src.d_pad, gi - src.off[part]);
    if (dedup_thr > -1.0f) {
      for (int j = 0; j < kept && !dup; ++j) {
        const CosAcc a = warp_cos_acc(s_rows + static_cast<size_t>(j) * d, row, d);
        dup = cos_from(a) > dedup_thr;
      }
      if (dup) continue;
    }
    const CosAcc a = warp_cos_acc(s_img, row, d);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      sims[kept] = cos_from(a);
      kept_idx[kept] = gi;
    }
    __syncwarp();
    ++kept;
  }
  return kept;
}

__global__ void consistency_emb_kernel(const tvc_detector_params p, long long nq, int d,
                                       const ConsistencyEmbArgs a, float* __restrict__ scores,
                                       uint8_t* __restrict__ flags, int rows_cap) {
  extern __shared__ __align__(16) float s_dyn[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  // per warp: image row + rows_cap kept rows, then small lists
  const size_t row_floats = (static_cast<size_t>(1 + rows_cap) * d + 3) & ~static_cast<size_t>(3);
  float* s_img = s_dyn + static_cast<size_t>(w) * row_floats;
  float* s_rows = s_img + d;
  float* s_lists = s_dyn + static_cast<size_t>(warps) * row_floats;
  constexpr int kXMax = TVC_MAX_VARIANTS * (TVC_MAX_VARIANTS - 1) / 2;
  constexpr int kListFloats = TVC_MAX_VARIANTS + 2 * TVC_MAX_REFS + kXMax + TVC_NSCORES;
  float* l_sv = s_lists + static_cast<size_t>(w) * (kListFloats + 2 * TVC_MAX_REFS);
  float* l_sr = l_sv + TVC_MAX_VARIANTS;
  float* l_sg = l_sr + TVC_MAX_REFS;
  float* l_sx = l_sg + TVC_MAX_REFS;
  float* l_out = l_sx + kXMax;
  long long* l_kept = reinterpret_cast<long long*>(l_out + TVC_NSCORES);  // TVC_MAX_REFS entries

  const int V = p.n_variants, R = p.n_retrieval, G = p.n_generative;
  for (long long q = static_cast<long long>(blockIdx.x) * warps + w; q < nq;
       q += static_cast<long long>(gridDim.x) * warps) {
    // image row -> smem
    load_row_to_smem(s_img, a.img, nullptr, d, d, q);
    float s0 = 0.f;
    {
      const CosAcc c = warp_cos_acc(s_img, a.txt + q * d, d);
      s0 = cos_from(c);
    }
    const float* var_q = a.var ? a.var + q * V * d : nullptr;
    const int nv = var_q ? V : 0;
    for (int v = 0; v < nv; ++v) {
      const CosAcc c = warp_cos_acc(s_img, var_q + static_cast<size_t>(v) * d, d);
      if (lane == 0) l_sv[v] = cos_from(c);
    }
    int nx = 0;
    for (int i = 0; i < nv; ++i)
      for (int j = i + 1; j < nv; ++j) {
        const CosAcc c = warp_cos_acc(var_q + static_cast<size_t>(i) * d,
                                      var_q + static_cast<size_t>(j) * d, d);
        if (lane == 0) l_sx[nx] = cos_from(c);
        ++nx;
      }
    int nr = 0;
    if (a.ret_idx && a.ret.nparts > 0)
      nr = select_refs(s_img, s_rows, d, a.ret, reinterpret_cast<const long long*>(a.ret_idx) + q * a.n_ret_cand,
                       a.n_ret_cand, min(R, rows_cap), p.dedup_threshold, l_sr, l_kept);
    int ng = 0;
    if (a.gen) {
      ng = a.g_cnt ? max(0, min(G, a.g_cnt[q])) : G;
      for (int g = 0; g < ng; ++g) {
        const CosAcc c = warp_cos_acc(s_img, a.gen + (q * G + g) * d, d);
        if (lane == 0) l_sg[g] = cos_from(c);
      }
    } else if (a.gen_idx && a.genr.nparts > 0) {
      ng = select_refs(s_img, s_rows, d, a.genr, reinterpret_cast<const long long*>(a.gen_idx) + q * a.n_gen_cand,
                       a.n_gen_cand, min(G, rows_cap), p.dedup_threshold, l_sg, l_kept);
    }
    __syncwarp();
    if (lane == 0) {
      uint8_t flag;
      finish_scores(p, s0, l_sv, nv, l_sr, nr, l_sg, ng, l_sx, nx, l_out, &flag);
      flags[q] = flag;
    }
    __syncwarp();
    if (lane < TVC_NSCORES) scores[q * TVC_NSCORES + lane] = l_out[lane];
    if (a.out_sv && lane < V) a.out_sv[q * V + lane] = lane < nv ? l_sv[lane] : 0.f;
    if (a.out_sr && lane < R) a.out_sr[q * R + lane] = lane < nr ? l_sr[lane] : 0.f;
    if (a.out_sg && lane < G) a.out_sg[q * G + lane] = lane < ng ? l_sg[lane] : 0.f;
    __syncwarp();
  }
}

}  // namespace

// =============================================================================== launchers
cudaError_t launch_prep_rows(const void* rows, int dtype, int64_t n, int d, int d_pad, bool normalize,
                             __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  long long blocks = (n * 32 + block - 1) / block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int grid = static_cast<int>(blocks);
  switch (dtype) {
    case TVC_F32:
      prep_rows_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(rows), n, d, d_pad,
                                                          normalize, out_bf16, out_f32);
      break;
    case TVC_BF16:
      prep_rows_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(rows), n, d, d_pad, normalize, out_bf16, out_f32);
      break;
    case TVC_F16:
      prep_rows_kernel<__half><<<grid, block, 0, stream>>>(static_cast<const __half*>(rows), n, d,
                                                           d_pad, normalize, out_bf16, out_f32);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_rerank(const float* cand_val, const int32_t* cand_idx, int64_t m, int splits,
                          int kp, int k, const float* q_f32, const float* g_f32, int d,
                          float threshold, int64_t global_row_offset, float* out_sim,
                          int64_t* out_idx, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  if (kp > kMaxKp) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((m + kRerankWarps - 1) / kRerankWarps);
  rerank_kernel<<<grid, kRerankWarps * 32, 0, stream>>>(
      cand_val, cand_idx, m, splits, kp, k, q_f32, g_f32, d, threshold, global_row_offset, out_sim,
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

cudaError_t launch_k_occurrence(const int64_t* idx, int64_t m, int k, int64_t idx_base,
                                int64_t n_bins, int32_t* counts, int sm_count, int* flag_scratch,
                                cudaStream_t stream) {
  const long long total = m * k;
  if (total <= 0 || n_bins <= 0) return cudaSuccess;
  const int block = 256;
  const long long quads = (total + 3) / 4 + 1;
  long long blocks = (quads + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count) * 16;
  if (blocks > cap) blocks = cap;
  const size_t hist_bytes = static_cast<size_t>(n_bins) * 4;
  // private shared-memory histograms pay off when each block sees many increments per bin flush
  const bool use_smem = hist_bytes <= 48 * 1024 && total >= n_bins * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(idx) & 15u) == 0;
  const long long* ip = reinterpret_cast<const long long*>(idx);
  const int g = static_cast<int>(blocks);
  // hot-bin detector: a flag word (per-stream scratch) written by the sampling pass, read by the main kernel
  int* flag = nullptr;
  if (!use_smem && flag_scratch != nullptr) {
    flag = flag_scratch;
    k_occurrence_sample_kernel<<<1, 1024, 0, stream>>>(ip, total, flag);
    note_launch();
  }
  if (use_smem) {
    if (aligned)
      k_occurrence_kernel<true, true><<<g, block, hist_bytes, stream>>>(ip, total, idx_base, n_bins, counts, flag);
    else
      k_occurrence_kernel<true, false><<<g, block, hist_bytes, stream>>>(ip, total, idx_base, n_bins, counts, flag);
  } else {
    if (aligned)
      k_occurrence_kernel<false, true><<<g, block, 0, stream>>>(ip, total, idx_base, n_bins, counts, flag);
    else
      k_occurrence_kernel<false, false><<<g, block, 0, stream>>>(ip, total, idx_base, n_bins, counts, flag);
  }
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

cudaError_t launch_consistency_sims(const tvc_detector_params& p, int64_t q, const float* s0,
                                    const float* sv, const float* sr, const int32_t* r_cnt,
                                    const float* sg, const int32_t* g_cnt, const float* sxv,
                                    float* scores, uint8_t* flags, cudaStream_t stream) {
  if (q <= 0) return cudaSuccess;
  const int V = p.n_variants, R = p.n_retrieval, G = p.n_generative;
  const int X = sxv ? V * (V - 1) / 2 : 0;
  auto up4 = [](int x) { return (x + 3) & ~3; };
  const size_t floats = up4(kSimsBlock * V) + up4(kSimsBlock * R) + up4(kSimsBlock * G) +
                        up4(kSimsBlock * X) + 16 + kSimsBlock * TVC_NSCORES;
  const size_t smem = floats * 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(consistency_sims_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = static_cast<int>((q + kSimsBlock - 1) / kSimsBlock);
  consistency_sims_kernel<<<grid, kSimsBlock, smem, stream>>>(p, q, s0, sv, sr, r_cnt, sg, g_cnt, sxv,
                                                              scores, flags);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_consistency_emb(const tvc_detector_params& p, int64_t q, int d,
                                   const ConsistencyEmbArgs& a, float* scores, uint8_t* flags,
                                   cudaStream_t stream) {
  if (q <= 0) return cudaSuccess;
  int rows_cap = p.n_retrieval > p.n_generative ? p.n_retrieval : p.n_generative;
  if (rows_cap < 1) rows_cap = 1;
  constexpr int kXMax = TVC_MAX_VARIANTS * (TVC_MAX_VARIANTS - 1) / 2;
  constexpr int kListFloats = TVC_MAX_VARIANTS + 2 * TVC_MAX_REFS + kXMax + TVC_NSCORES;
  const size_t row_floats = (static_cast<size_t>(1 + rows_cap) * d + 3) & ~static_cast<size_t>(3);
  const size_t per_warp = (row_floats + kListFloats + 2 * TVC_MAX_REFS) * 4;
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  if (per_warp * warps > 220 * 1024) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(consistency_emb_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  long long blocks = (q + warps - 1) / warps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  consistency_emb_kernel<<<static_cast<int>(blocks), warps * 32, per_warp * warps, stream>>>(
      p, q, d, a, scores, flags, rows_cap);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tvc
