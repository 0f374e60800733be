// Kernel (b): the variant-consistency reduction.  Per query it turns similarities (or the embedding
// rows themselves) into the statistics and decisions of the reference's three scoring stacks:
//   AdversarialDetector            src/detector.py:441-590, 643-682, 399
//   MultiModalDefenseDetector      experiments/defenses/detector.py:228-325
//   ConsistencyChecker             experiments/defenses/consistency_checker.py:119-272
//   README sigma rule              README.md:474-482, 846
// Statistics are fp64 on fp32 similarities like the reference's Python floats -> np.mean/np.std.
//
// Both modes are HBM-bound by design:
//   * similarity-fed  - a block stages the contiguous similarity slabs of its queries through shared
//     memory with 128-bit loads, one thread reduces one query, results leave as coalesced stores;
//   * embedding-fed   - a persistent, warp-specialised CTA per SM.  One producer warp per stage of a
//     3-stage shared-memory ring walks the candidate lists and pulls the ~20 rows a query needs
//     (image, text, V variants, the first R distinct retrieval candidates, G generative rows) from
//     HBM - or a PEER GPU's HBM over NVLink - with cp.async.bulk (TMA) completing on mbarriers.
//     Sixteen consumer warps draw (query, task) units from one CTA-wide queue: a task is one resident
//     row A against up to 5 rows B, register-blocked, so the query's ~75 dot products cost ~20 units.
//     The warp that draws a query's last unit waits for the rest, runs the greedy de-duplication on
//     the resulting cosine matrix (one lane per candidate), releases the stage and writes the
//     similarity lists; the similarity-fed kernel then turns them into statistics and decisions.
//     Measured limits (scripts/trace_emb.py): the kernel is bound by per-stage latency (index load ->
//     bulk copies -> tasks -> selection ~ 9 us) over the 3 stages that fit in 227 KB, not by HBM.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "tvc_internal.h"
#include "tvc_ptx.cuh"

namespace tvc {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ------------------------------------------------------------------------------- fp64 statistics
// Division of a double by a small positive integer with a tabulated correctly rounded reciprocal and
// two FMA-residual refinements (Markstein): the result is the correctly rounded quotient, i.e. what
// IEEE division (np.mean / np.var) returns, for a fraction of the cost of the DP division sequence.
constexpr int kRcpN = 128;   // covers every count: V, R, G <= 16, V(V-1)/2 <= 120, R+G <= 32

// 1/i, correctly rounded by the host compiler (IEEE division of constants).  Round 1 had every block divide the
// table out itself: 128 DP divisions = 3.7 % of all instructions of the statistics kernel (ncu source view).
#define TVC_RCP1(i) ((i) == 0 ? 0.0 : 1.0 / static_cast<double>(i))
#define TVC_RCP8(b) TVC_RCP1(b), TVC_RCP1(b + 1), TVC_RCP1(b + 2), TVC_RCP1(b + 3), TVC_RCP1(b + 4), TVC_RCP1(b + 5), \
                    TVC_RCP1(b + 6), TVC_RCP1(b + 7)
#define TVC_RCP64(b) TVC_RCP8(b), TVC_RCP8(b + 8), TVC_RCP8(b + 16), TVC_RCP8(b + 24), TVC_RCP8(b + 32), \
                     TVC_RCP8(b + 40), TVC_RCP8(b + 48), TVC_RCP8(b + 56)
__constant__ double c_rcp[kRcpN] = {TVC_RCP64(0), TVC_RCP64(64)};

__device__ __forceinline__ void fill_rcp_table(double* rcp, int tid, int nthreads) {
  for (int i = tid; i < kRcpN; i += nthreads) rcp[i] = c_rcp[i];
}

// sqrt of a finite double >= 0, correctly rounded like IEEE sqrt (np.std / math.sqrt), in 15 instructions
// instead of the ~40 of the library sequence with its special-case branches: a 22-bit reciprocal square root
// from the fp32 unit seeds two coupled Newton steps (g -> sqrt x, h -> 1 / (2 sqrt x)), and the FMA residual
// step g + h (x - g g) rounds correctly (Markstein).  Checked against IEEE sqrt on 3e5 values with an exact-FMA
// emulation, seed perturbed by +-3 fp32 ulp: no mismatch.  Values the fp32 seed cannot represent take the
// library path.
__device__ __forceinline__ double sqrt_lean(double x) {
  if (!(x > 1e-30 && x < 1e30)) return sqrt(x);
  const double r = static_cast<double>(rsqrtf(static_cast<float>(x)));
  double g = x * r, h = 0.5 * r;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
  }
  return fma(fma(-g, g, x), h, g);
}

__device__ __forceinline__ double div_n(double s, int n, const double* rcp) {
  const double r = rcp[n], dn = static_cast<double>(n);
  double q = s * r;
  q = fma(fma(-q, dn, s), r, q);
  q = fma(fma(-q, dn, s), r, q);
  return q;
}

struct Stats {
  double sum, mean, ss, var, sd;   // ss = sum of squared deviations, var = ss / n (ddof = 0)
  float mn, mx;
};

// x[0..n) in shared memory.  Two passes that RE-READ the values (an LDS + a convert per element) instead of
// keeping them in registers: the converted values of the four groups were 28 fp64 registers under a 64-register
// cap, and with a run-time count every unrolled element carried its own predicate (ncu / SASS of round 1: 1830
// instructions per query, a third of them ISETP / FSEL / IMAD.MOV).  FIXED > 0 = the count is the compile-time
// constant FIXED (straight-line code, no predicates); FIXED == 0 = run-time count, a plain loop.  The order of
// the additions is the same in every form, so all kernels that share this function agree to the bit.
template <int FIXED>
__device__ __forceinline__ Stats stats_of(const float* x, int n, const double* rcp, bool want_sd) {
  Stats s{0., 0., 0., 0., 0., 0.f, 0.f};
  if (n <= 0) return s;
  float mn = x[0], mx = x[0];
  double sum = 0., ss = 0.;
  if constexpr (FIXED > 0) {
#pragma unroll
    for (int i = 0; i < FIXED; ++i) {
      const float f = x[i];
      mn = fminf(mn, f);
      mx = fmaxf(mx, f);
      sum += static_cast<double>(f);
    }
    s.mean = div_n(sum, FIXED, rcp);
#pragma unroll
    for (int i = 0; i < FIXED; ++i) {
      const double dlt = static_cast<double>(x[i]) - s.mean;
      ss = fma(dlt, dlt, ss);
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
      const float f = x[i];
      mn = fminf(mn, f);
      mx = fmaxf(mx, f);
      sum += static_cast<double>(f);
    }
    s.mean = div_n(sum, n, rcp);
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
      const double dlt = static_cast<double>(x[i]) - s.mean;
      ss = fma(dlt, dlt, ss);
    }
  }
  s.sum = sum;
  s.ss = ss;
  s.var = div_n(ss, n, rcp);
  s.sd = want_sd ? sqrt_lean(s.var) : 0.;
  s.mn = mn;
  s.mx = mx;
  return s;
}

// the counts the defaults produce (V = 5, X = 10, G = 3) take the straight-line form
template <int A, int B>
__device__ __forceinline__ Stats stats_dispatch(const float* x, int n, const double* rcp, bool want_sd) {
  if (n == A) return stats_of<A>(x, n, rcp, want_sd);
  if (B > 0 && n == B) return stats_of<(B > 0 ? B : 1)>(x, n, rcp, want_sd);
  return stats_of<0>(x, n, rcp, want_sd);
}

__device__ __forceinline__ double clipd(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

// Scores and decisions of one query from the four groups' statistics (text variants, retrieval
// references, generative references, variant pairs).
__device__ __forceinline__ void combine_scores(const tvc_detector_params& p, const double* rcp, float s0f,
                                               const Stats& tv, int nv, const Stats& rt, int nr,
                                               const Stats& gn, int ng, float* out,
                                               int out_stride, uint8_t* flag) {
  const double s0 = s0f;

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
#pragma unroll
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
        agg = div_n(sum, cnt, rcp);
    }
  }
  const bool det_adv = agg > static_cast<double>(p.detection_threshold);

  // --- MultiModalDefenseDetector scores (experiments/defenses/detector.py:228-300)
  const double tv_c = nv > 0 ? tv.mean : s0;
  const double tv_s = nv > 0 ? tv.sd : 0.0;
  const double rt_c = nr > 0 ? rt.mean : 0.0, rt_s = nr > 0 ? rt.sd : 0.0;
  const double gn_c = ng > 0 ? gn.mean : 0.0, gn_s = ng > 0 ? gn.sd : 0.0;
  const double four[4] = {s0, tv_c, rt_c, gn_c};
  int nvalid = 0;
  double vsum = 0.;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (four[i] > 0.) {
      vsum += four[i];
      ++nvalid;
    }
  double vmean = 0., vvar = 0.;
  if (nvalid > 0) {
    vmean = div_n(vsum, nvalid, rcp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (four[i] > 0.) vvar = fma(four[i] - vmean, four[i] - vmean, vvar);
    vvar = div_n(vvar, nvalid, rcp);
  }
  const double cmv = nvalid < 2 ? 0.0 : vvar;

  // --- ConsistencyChecker (experiments/defenses/consistency_checker.py:119-272)
  double overall = 0.0;
  if (p.voting == 0) {
    overall = nvalid > 0 ? vmean : 0.0;
  } else {
    double w[4];
    if (p.voting == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = p.cc_weights[i];
    } else {
      w[0] = 1.0;
      w[1] = 1.0 / (1.0 + tv_s);
      w[2] = 1.0 / (1.0 + rt_s);
      w[3] = 1.0 / (1.0 + gn_s);
      const double t = w[0] + w[1] + w[2] + w[3];
      if (t > 0.) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] /= t;
      }
    }
    double ws = 0., tw = 0.;
#pragma unroll
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
    const double avg_std = div_n(tv_s + rt_s + gn_s, 3, rcp);
    if (avg_std > 0.2) thr += 0.05;
    thr = clipd(thr, 0.1, 0.9);
  }
  const bool cc_adv = overall < thr;
  const double dist_conf = fabs(overall - thr) / thr;
  const double cons_conf = nvalid > 1 ? 1.0 - sqrt_lean(vvar) : 0.5;
  const double var_conf = 1.0 - (cmv < 1.0 ? cmv : 1.0);
  const double conf = clipd(div_n(dist_conf + cons_conf + var_conf, 3, rcp), 0.0, 1.0);

  // --- README sigma rule over all references (README.md:474-482, 846): population std of the
  // retrieval and generative similarities together, from the two groups' moments
  double sigma = 0.0;
  {
    const int n = nr + ng;
    if (n > 0) {
      const double mu = div_n(rt.sum + gn.sum, n, rcp);
      double acc = rt.ss + gn.ss;
      acc = fma(static_cast<double>(nr) * (rt.mean - mu), rt.mean - mu, acc);
      acc = fma(static_cast<double>(ng) * (gn.mean - mu), gn.mean - mu, acc);
      sigma = sqrt_lean(div_n(acc, n, rcp));
    }
  }
  const bool sig_adv = sigma > static_cast<double>(p.sigma_threshold);

  auto put = [&](int col, double v) { out[col * out_stride] = static_cast<float>(v); };
  out[TVC_S_ORIGINAL * out_stride] = s0f;
  put(TVC_S_TV_MEAN, tv_c);
  put(TVC_S_TV_STD, tv_s);
  out[TVC_S_TV_MIN * out_stride] = nv > 0 ? tv.mn : s0f;
  put(TVC_S_TV_VAR, nv > 0 ? tv.var : 0.0);
  put(TVC_S_RET_MEAN, rt_c);
  put(TVC_S_RET_STD, rt_s);
  put(TVC_S_GEN_MEAN, gn_c);
  put(TVC_S_GEN_STD, gn_s);
  out[TVC_S_GEN_MAX * out_stride] = ng > 0 ? gn.mx : 0.f;
  put(TVC_S_CROSS_MODAL_VAR, cmv);
  put(TVC_S_DET_TV, det_tv);
  put(TVC_S_DET_SD, det_sd);
  put(TVC_S_DET_C, det_c);
  put(TVC_S_DET_AGG, agg);
  put(TVC_S_CC_OVERALL, overall);
  put(TVC_S_CC_THRESHOLD, thr);
  put(TVC_S_CC_CONFIDENCE, conf);
  out[TVC_S_N_RET * out_stride] = static_cast<float>(nr);
  out[TVC_S_N_GEN * out_stride] = static_cast<float>(ng);
  put(TVC_S_REF_SIGMA, sigma);
  *flag = static_cast<uint8_t>((det_adv ? TVC_FLAG_DET_ADV : 0u) | (cc_adv ? TVC_FLAG_CC_ADV : 0u) |
                               (sig_adv ? TVC_FLAG_SIGMA_ADV : 0u));
}

// sv/sr/sg/sx point at this query's similarity lists (shared memory)
__device__ __forceinline__ void finish_scores_any(const tvc_detector_params& p, const double* rcp, float s0,
                                                  const float* sv, int nv, const float* sr, int nr,
                                                  const float* sg, int ng, const float* sx, int nx,
                                                  float* out, int out_stride, uint8_t* flag) {
  {   // the variant<->variant columns depend on nothing else: written first, so their moments are not kept live
    const Stats xv = stats_dispatch<10, 0>(sx, nx, rcp, false);
    out[TVC_S_XV_MEAN * out_stride] = static_cast<float>(xv.mean);
    out[TVC_S_XV_MIN * out_stride] = xv.mn;
    out[TVC_S_XV_VAR * out_stride] = static_cast<float>(xv.var);
  }
  const Stats tv = stats_dispatch<5, 0>(sv, nv, rcp, true);
  const Stats rt = stats_dispatch<10, 5>(sr, nr, rcp, true);
  const Stats gn = stats_dispatch<3, 0>(sg, ng, rcp, true);
  combine_scores(p, rcp, s0, tv, nv, rt, nr, gn, ng, out, out_stride, flag);
}

// ------------------------------------------------------------------------------- similarity-fed
constexpr int kSimsBlock = 128;
constexpr int kSimsMinBlocks = 8;
// Result tile [query][28]: 24 scores + 4 floats of padding keep every row 16-byte aligned, so the thread that
// computed a query copies its own row out with 128-bit accesses - no barrier, no index arithmetic (the copy-out
// loop over the whole tile was 8 % of the kernel's instructions).  Scalar stores into the tile are 4-way bank
// conflicted (stride 28), 21 of them per query.
constexpr int kOutStride = TVC_NSCORES + 4;

// global -> shared bulk copy (TMA, no tensor map), completion bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void stage_slab(float* dst, const float* __restrict__ src, long long q0,
                                           int nq, int width) {
  if (src == nullptr || width == 0) return;
  const long long base = q0 * width;
  const int total = nq * width;
  const float* s = src + base;
  if (((reinterpret_cast<uintptr_t>(s) & 15u) == 0)) {
    const int n4 = total >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      reinterpret_cast<float4*>(dst)[i] = __ldcs(reinterpret_cast<const float4*>(s) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) dst[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = s[i];
  }
}

template <int MINB>   // resident blocks per SM the register budget is cut for: 8 -> 64 registers, 6 -> 80, 4 -> 128
__global__ void __launch_bounds__(kSimsBlock, MINB)
consistency_sims_kernel(const tvc_detector_params p, long long nq_total, const float* __restrict__ s0,
                        const float* __restrict__ sv, const float* __restrict__ sr,
                        const int32_t* __restrict__ r_cnt, const float* __restrict__ sg,
                        const int32_t* __restrict__ g_cnt, const float* __restrict__ sxv,
                        float* __restrict__ scores, uint8_t* __restrict__ flags) {
  extern __shared__ __align__(16) float s_buf[];
  __shared__ double s_rcp[kRcpN];
  __shared__ __align__(8) uint64_t s_bar;
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
  // The block's four similarity slabs are contiguous in global memory: one thread hands them to the copy
  // engine (cp.async.bulk completing on an mbarrier) instead of 128 threads looping over float4s - that loop was
  // 13 % of the kernel's instructions and its load waits a quarter of all stall samples (ncu source view).
  // Needs 16-byte aligned sources and sizes, which holds for every block but a ragged last one.
  auto bulk_ok = [&](const float* src, int width) {
    return src == nullptr || width == 0 ||
           (((reinterpret_cast<uintptr_t>(src + q0 * width) & 15u) == 0) && ((nq * width) & 3) == 0);
  };
  const bool bulk = bulk_ok(sv, V) && bulk_ok(sr, R) && bulk_ok(sg, G) && bulk_ok(sxv, X);   // block-uniform
  if (bulk && threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    uint32_t bytes = 0;
    auto slab_bytes = [&](const float* src, int width) { return (src && width) ? static_cast<uint32_t>(nq * width * 4) : 0u; };
    bytes = slab_bytes(sv, V) + slab_bytes(sr, R) + slab_bytes(sg, G) + slab_bytes(sxv, X);
    mbar_arrive_expect_tx(&s_bar, bytes);
    if (slab_bytes(sv, V)) bulk_g2s(b_sv, sv + q0 * V, slab_bytes(sv, V), &s_bar);
    if (slab_bytes(sr, R)) bulk_g2s(b_sr, sr + q0 * R, slab_bytes(sr, R), &s_bar);
    if (slab_bytes(sg, G)) bulk_g2s(b_sg, sg + q0 * G, slab_bytes(sg, G), &s_bar);
    if (slab_bytes(sxv, X)) bulk_g2s(b_sx, sxv + q0 * X, slab_bytes(sxv, X), &s_bar);
  }
  fill_rcp_table(s_rcp, threadIdx.x, blockDim.x);
  if (!bulk) {
    stage_slab(b_sv, sv, q0, nq, V);
    stage_slab(b_sr, sr, q0, nq, R);
    stage_slab(b_sg, sg, q0, nq, G);
    stage_slab(b_sx, sxv, q0, nq, X);
  }
  const int t = threadIdx.x;
  // this thread's scalars travel while the slabs do
  const long long q = q0 + t;
  float my_s0 = 0.f;
  int nr = 0, ng = 0;
  if (t < nq) {
    my_s0 = s0[q];
    nr = sr ? (r_cnt ? max(0, min(R, r_cnt[q])) : R) : 0;
    ng = sg ? (g_cnt ? max(0, min(G, g_cnt[q])) : G) : 0;
  }
  __syncthreads();   // barrier initialised (bulk) / slabs staged (fallback), reciprocal table filled
  if (bulk) mbar_wait(&s_bar, 0);
  if (t < nq) {
    const int nv = sv ? V : 0;
    uint8_t flag = 0;
    float* row = b_out + t * kOutStride;
    finish_scores_any(p, s_rcp, my_s0, b_sv + t * V, nv, b_sr + t * R, nr, b_sg + t * G, ng, b_sx + t * X, X, row, 1,
                      &flag);
    flags[q] = flag;
    // own row out: TVC_NSCORES is a multiple of 4 and the tile rows are 16-byte aligned
    static_assert(TVC_NSCORES % 4 == 0 && kOutStride % 4 == 0, "float4 copy-out");
    float* dst = scores + q * TVC_NSCORES;
    if ((reinterpret_cast<uintptr_t>(scores) & 15u) == 0) {
#pragma unroll
      for (int c = 0; c < TVC_NSCORES / 4; ++c)
        __stcs(reinterpret_cast<float4*>(dst) + c, reinterpret_cast<const float4*>(row)[c]);
    } else {
#pragma unroll
      for (int c = 0; c < TVC_NSCORES; ++c) dst[c] = row[c];
    }
  }
}

// ------------------------------------------------------------------------------- embedding-fed
// torch.cosine_similarity: x.y / max(|x||y|, eps), eps = 1e-8
__device__ __forceinline__ float cos_of(float dot, float na, float nb) {
  return dot / fmaxf(sqrtf(na) * sqrtf(nb), 1e-8f);
}

struct CosAcc {
  float dot, na, nb;
};
// one warp, rows anywhere (generic loads); used by the generic kernel and by the finisher's slow path
__device__ __forceinline__ CosAcc warp_cos_acc(const float* __restrict__ a, const float* __restrict__ b, int d) {
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
__device__ __forceinline__ float cos_from(const CosAcc& c) { return cos_of(c.dot, c.na, c.nb); }

__device__ __forceinline__ int find_part(const RowSource& src, long long gi) {
  int part = -1;
  for (int p = 0; p < src.nparts; ++p)
    if (gi >= src.off[p] && gi < src.off[p] + src.n[p]) part = p;
  return part;
}

__device__ __forceinline__ void load_row_to_smem(float* dst, const float* g_f32, const __nv_bfloat16* g_bf16,
                                                 int d, int d_pad, long long gi) {
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

// ---- generic kernel (any d / alignment, bf16-only galleries): one warp per query -----------------
// Greedy reference selection (experiments/defenses/detector.py:184-204, 302-325): walk the
// candidate list in order, drop repeated indices and rows whose cosine to an already kept row
// exceeds dedup_threshold, stop at `cap` kept rows; sims[j] = cos(image, kept row j).
__device__ int select_refs(const float* s_img, float* s_rows, int d, const RowSource& src,
                           const long long* cand, int ncand, int cap, float dedup_thr, float* sims,
                           long long* kept_idx) {
  int kept = 0;
  for (int c = 0; c < ncand && kept < cap; ++c) {
    const long long gi = cand[c];
    const int part = find_part(src, gi);
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

constexpr int kXMax = TVC_MAX_VARIANTS * (TVC_MAX_VARIANTS - 1) / 2;
constexpr int kListFloats = TVC_MAX_VARIANTS + 2 * TVC_MAX_REFS + kXMax + TVC_NSCORES;

__global__ void consistency_emb_generic_kernel(const tvc_detector_params p, long long nq, int d,
                                               const ConsistencyEmbArgs a, float* __restrict__ scores,
                                               uint8_t* __restrict__ flags, int rows_cap) {
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ double s_rcp[kRcpN];
  fill_rcp_table(s_rcp, threadIdx.x, blockDim.x);
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  // per warp: image row + rows_cap kept rows, then small lists
  const size_t row_floats = (static_cast<size_t>(1 + rows_cap) * d + 3) & ~static_cast<size_t>(3);
  float* s_img = s_dyn + static_cast<size_t>(w) * row_floats;
  float* s_rows = s_img + d;
  float* s_lists = s_dyn + static_cast<size_t>(warps) * row_floats;
  float* l_sv = s_lists + static_cast<size_t>(w) * (kListFloats + 2 * TVC_MAX_REFS);
  float* l_sr = l_sv + TVC_MAX_VARIANTS;
  float* l_sg = l_sr + TVC_MAX_REFS;
  float* l_sx = l_sg + TVC_MAX_REFS;
  float* l_out = l_sx + kXMax;
  long long* l_kept = reinterpret_cast<long long*>(l_out + TVC_NSCORES);  // TVC_MAX_REFS entries

  const int V = p.n_variants, R = p.n_retrieval, G = p.n_generative;
  for (long long q = static_cast<long long>(blockIdx.x) * warps + w; q < nq;
       q += static_cast<long long>(gridDim.x) * warps) {
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
      finish_scores_any(p, s_rcp, s0, l_sv, nv, l_sr, nr, l_sg, ng, l_sx, nx, l_out, 1, &flag);
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

// ---- pipelined kernel -----------------------------------------------------------------------------
constexpr int kEmbConsumers = 16;                       // consumer warps
constexpr int kEmbMaxStages = 4;
constexpr int kEmbProducers = kEmbMaxStages;            // producer warps: one per stage of the ring
constexpr int kEmbThreads = (kEmbConsumers + kEmbProducers) * 32;
constexpr int kJB = 5;                                  // B rows per task (register block)
constexpr int kMaxTasks = 224;
constexpr int kMaxStageRows = 2 + TVC_MAX_VARIANTS + 2 * TVC_MAX_REFS;   // 50

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// optional tracing of block 0 (a.trace != null): 8 timestamps per query, see scripts/trace_emb.py
#define TVC_TRACE(i, k)                                                                              \
  do {                                                                                               \
    if (a.trace != nullptr && blockIdx.x == 0 && (i) < 512) atomicMin(a.trace + (i)*8 + (k), gtime_ns()); \
  } while (0)

// One task = dots of row A against rows B0 .. B0+nb-1 of the stage (+ squared norms for the image
// group).  Groups: 0 = image/text/variants (always resident), 1 = retrieval rows, 2 = generative rows.
struct EmbTask {
  short a_row, b_row0;      // stage row numbers
  unsigned char nb, norms;  // B rows; 1 = image task (also writes |A|^2, |B|^2), 0 / 2 = off-diagonal /
                            // diagonal block of a group's pair table (`pad` A rows x nb B rows)
  unsigned char ga, gb;     // validity groups of A and of the B rows
  short ia, ib0;            // index of A / first B inside its group
  short out;                // result offset of (A, B0); consecutive B -> consecutive floats
  short pad;                // A rows of a block task
};

// per-stage control block and results (shared memory)
struct EmbStage {
  long long q;
  long long pf_idx[2][TVC_MAX_REFS];   // global indices of the prefetched rows (0 retrieval, 1 generative)
  int pf_n[2];                         // rows prefetched
  int pf_end[2];                       // candidate position after the last prefetched one
  int n_gen_direct;                    // valid direct generative rows (g_cnt)
  // results: squared norms and image dots per stage row, pair dots per group
  float nrm2[kMaxStageRows];
  float d_img[kMaxStageRows];
  float d_var[TVC_MAX_VARIANTS * TVC_MAX_VARIANTS];
  float d_ref[2][TVC_MAX_REFS * TVC_MAX_REFS];
};

// finisher scratch (per consumer warp)
struct EmbLists {
  float sv[TVC_MAX_VARIANTS];
  float sr[TVC_MAX_REFS];
  float sg[TVC_MAX_REFS];
  float sx[kXMax];
  float sq[kMaxStageRows];   // |row| of every stage row
  float dimg[kMaxStageRows]; // image . row (copied out so that the stage can be released early)
  long long kept_idx[TVC_MAX_REFS];
  int kept_slot[TVC_MAX_REFS];
  int info[4];   // kept, need_slow
};

// Sums N per-lane values over the warp with N-1 + log2(32/N) shuffles instead of 5 N: at every level
// half of the values go to the partner lane, which is the same addition tree as the xor butterfly
// (bit-identical results).  Returns the total of value number bitrev-ish index `multi_index<N>(lane)`.
template <int N>
__device__ __forceinline__ float warp_sum_multi(float (&v)[N], int lane) {
  static_assert(N == 8 || N == 16 || N == 32, "N");
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[2 * i] : v[2 * i + 1];
      const float keep = upper ? v[2 * i + 1] : v[2 * i];
      v[i] = keep + __shfl_xor_sync(kFull, send, off);
    }
  }
  float r = v[0];
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(kFull, r, off);
  return r;
}
template <int N>
__device__ __forceinline__ int multi_index(int lane) {
  if (N == 32)
    return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2) | (((lane >> 1) & 1) << 3) |
           ((lane & 1) << 4);
  if (N == 16) return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2) | (((lane >> 1) & 1) << 3);
  return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2);
}

template <bool NORMS, int NB>
__device__ __forceinline__ void run_task_nb(const EmbTask& t, const float* rows, int d, EmbStage* st, float* out) {
  const int lane = threadIdx.x & 31;
  const float4* A = reinterpret_cast<const float4*>(rows + static_cast<size_t>(t.a_row) * d);
  const float4* B = reinterpret_cast<const float4*>(rows + static_cast<size_t>(t.b_row0) * d);
  const int d4 = d >> 2;
  constexpr int NV = NORMS ? 16 : 8;     // dot[0..4] | nbn[0..4] at 5..9, |A|^2 at 10
  float v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = 0.f;
#pragma unroll 2
  for (int c = lane; c < d4; c += 32) {
    const float4 x = A[c];
    float4 y[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) y[j] = B[static_cast<size_t>(j) * d4 + c];
    // component-major order: NB (or 2 NB + 1) independent chains advance together
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = fmaf(x.x, y[j].x, v[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = fmaf(x.y, y[j].y, v[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = fmaf(x.z, y[j].z, v[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = fmaf(x.w, y[j].w, v[j]);
    if (NORMS) {
      v[10] = fmaf(x.x, x.x, v[10]);
#pragma unroll
      for (int j = 0; j < NB; ++j) v[5 + j] = fmaf(y[j].x, y[j].x, v[5 + j]);
      v[10] = fmaf(x.y, x.y, v[10]);
#pragma unroll
      for (int j = 0; j < NB; ++j) v[5 + j] = fmaf(y[j].y, y[j].y, v[5 + j]);
      v[10] = fmaf(x.z, x.z, v[10]);
#pragma unroll
      for (int j = 0; j < NB; ++j) v[5 + j] = fmaf(y[j].z, y[j].z, v[5 + j]);
      v[10] = fmaf(x.w, x.w, v[10]);
#pragma unroll
      for (int j = 0; j < NB; ++j) v[5 + j] = fmaf(y[j].w, y[j].w, v[5 + j]);
    }
  }
  const float r = warp_sum_multi<NV>(v, lane);
  const int idx = multi_index<NV>(lane);
  if ((lane & (NORMS ? 1 : 3)) == 0) {   // one lane per value
    if (idx < kJB) {
      if (idx < NB) out[t.out + idx] = r;
    } else if (NORMS) {
      if (idx < 2 * kJB) {
        if (idx - kJB < NB) st->nrm2[t.b_row0 + idx - kJB] = r;
      } else if (idx == 10) {
        st->nrm2[t.a_row] = r;
      }
    }
  }
}

// Block task: the pair dots of up to 5 rows A against up to 5 rows B of one group (DIAG: A == B, pairs
// i < j only).  25 accumulators per lane against 10 row reads - the tasks are bound by shared-memory
// bandwidth (128 B/clk/SM), and one-row-against-five tasks read 55 rows for the 45 pairs of 10
// references where three block tasks read 20.  Per pair the additions happen in the same order as
// everywhere else (lane-strided float4s, then the butterfly tree), so results stay bit-identical.
template <bool DIAG>
__device__ __forceinline__ void run_block_task(const EmbTask& t, int na, int nb, const float* rows, int d,
                                               float* out) {
  const int lane = threadIdx.x & 31;
  const float4* A = reinterpret_cast<const float4*>(rows + static_cast<size_t>(t.a_row) * d);
  const float4* B = reinterpret_cast<const float4*>(rows + static_cast<size_t>(t.b_row0) * d);
  const int d4 = d >> 2;
  float v[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = 0.f;
  for (int c = lane; c < d4; c += 32) {
    float4 x[kJB];
#pragma unroll
    for (int i = 0; i < kJB; ++i) x[i] = i < na ? A[static_cast<size_t>(i) * d4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kJB; ++j) {
      if (j >= nb) continue;
      const float4 y = DIAG ? x[j] : B[static_cast<size_t>(j) * d4 + c];
      // component-major: the (up to) five chains of column j advance together, so no FMA waits on
      // the one issued just before it
#pragma unroll
      for (int i = 0; i < kJB; ++i)
        if (!DIAG || i < j) v[i * kJB + j] = fmaf(x[i].x, y.x, v[i * kJB + j]);
#pragma unroll
      for (int i = 0; i < kJB; ++i)
        if (!DIAG || i < j) v[i * kJB + j] = fmaf(x[i].y, y.y, v[i * kJB + j]);
#pragma unroll
      for (int i = 0; i < kJB; ++i)
        if (!DIAG || i < j) v[i * kJB + j] = fmaf(x[i].z, y.z, v[i * kJB + j]);
#pragma unroll
      for (int i = 0; i < kJB; ++i)
        if (!DIAG || i < j) v[i * kJB + j] = fmaf(x[i].w, y.w, v[i * kJB + j]);
    }
  }
  const float r = warp_sum_multi<32>(v, lane);
  const int idx = multi_index<32>(lane);
  const int i = idx / kJB, j = idx - i * kJB;
  if (idx < kJB * kJB && i < na && j < nb && (!DIAG || i < j))
    out[(t.ia + i) * TVC_MAX_REFS + t.ib0 + j] = r;
}

template <bool NORMS>
__device__ __forceinline__ void run_task(const EmbTask& t, int nb, const float* rows, int d, EmbStage* st,
                                         float* out) {
  static_assert(kJB == 5, "dispatch below covers 1..5");
  switch (nb) {
    case 5: run_task_nb<NORMS, 5>(t, rows, d, st, out); break;
    case 4: run_task_nb<NORMS, 4>(t, rows, d, st, out); break;
    case 3: run_task_nb<NORMS, 3>(t, rows, d, st, out); break;
    case 2: run_task_nb<NORMS, 2>(t, rows, d, st, out); break;
    default: run_task_nb<NORMS, 1>(t, rows, d, st, out); break;
  }
}

// Greedy selection over the prefetched rows using the pair dots of the stage (one lane per row, the
// order-dependent part resolved with shuffles); falls back to walking the rest of the candidate list
// with synchronous loads when de-duplication dropped a row.  Returns the number kept; sims[j] =
// cos(image, kept j).  Warp-cooperative (all 32 lanes call it).
__device__ int select_from_stage(const tvc_detector_params& p, EmbStage* st, int grp, float* rows, int d,
                                 int row_base /* stage row of the group's slot 0 */, const RowSource& src,
                                 const long long* cand, int ncand, int cap, float* sims, EmbLists* L) {
  const int lane = threadIdx.x & 31;
  const float thr = p.dedup_threshold;
  const int pf_n = st->pf_n[grp];
  const float* pd = st->d_ref[grp];
  // lane s owns prefetched row s: its cosine to the image and the set of earlier rows it duplicates
  unsigned dup_of = 0;
  float sim = 0.f;
  if (lane < pf_n) {
    const float ns = L->sq[row_base + lane];
    sim = st->d_img[row_base + lane] / fmaxf(L->sq[0] * ns, 1e-8f);
    if (thr > -1.0f)
      for (int j = 0; j < lane; ++j) {
        const float c = pd[j * TVC_MAX_REFS + lane] / fmaxf(L->sq[row_base + j] * ns, 1e-8f);
        if (c > thr) dup_of |= 1u << j;
      }
  }
  // greedy resolution in list order (every lane computes the same kept set)
  unsigned kept_mask = 0;
  int kept = 0;
  for (int s = 0; s < pf_n; ++s) {
    const unsigned m = __shfl_sync(kFull, dup_of, s);
    if (kept < cap && (m & kept_mask) == 0) {
      kept_mask |= 1u << s;
      ++kept;
    }
  }
  if (lane < pf_n && ((kept_mask >> lane) & 1u)) {
    const int pos = __popc(kept_mask & ((1u << lane) - 1u));
    sims[pos] = sim;
    L->kept_slot[pos] = lane;
    L->kept_idx[pos] = st->pf_idx[grp][lane];
  }
  __syncwarp();
  if (!(kept < cap && kept < pf_n && st->pf_end[grp] < ncand)) return kept;
  // slow path: a prefetched row was dropped, so later candidates may still qualify
  const float* s_img = rows;
  for (int c = st->pf_end[grp]; c < ncand && kept < cap; ++c) {
    const long long gi = cand[c];
    const int part = find_part(src, gi);
    if (part < 0) continue;
    bool dup = false;
    for (int j = 0; j < kept; ++j) dup |= (L->kept_idx[j] == gi);
    if (dup) continue;
    // a free slot: lowest slot not used by a kept row (kept < cap <= slots)
    unsigned used = 0;
    for (int j = 0; j < kept; ++j) used |= 1u << L->kept_slot[j];
    const int slot = __ffs(~used) - 1;
    float* row = rows + static_cast<size_t>(row_base + slot) * d;
    load_row_to_smem(row, src.f32[part], src.bf16[part], d, src.d_pad, gi - src.off[part]);
    if (thr > -1.0f) {
      for (int j = 0; j < kept && !dup; ++j) {
        const CosAcc a = warp_cos_acc(rows + static_cast<size_t>(row_base + L->kept_slot[j]) * d, row, d);
        dup = cos_from(a) > thr;
      }
      if (dup) continue;
    }
    const CosAcc a = warp_cos_acc(s_img, row, d);
    __syncwarp();
    if (lane == 0) {
      sims[kept] = cos_from(a);
      L->kept_idx[kept] = gi;
      L->kept_slot[kept] = slot;
    }
    __syncwarp();
    ++kept;
  }
  return kept;
}

// producer side: take the first `cap` distinct valid indices of a candidate list, start their copies.
// Which candidates of the first chunk are taken depends on the list alone, so that part (shard lookup, the votes
// and the 64-bit match) is planned BEFORE the producer waits for its stage - it was 1 of the 1.8 us between
// "stage free" and "copies issued" in the per-query trace (scripts/trace_emb.py), i.e. of every stage's cycle.
struct ChunkPlan {
  long long gi;
  int part;
  unsigned keep;   // lanes of the chunk whose candidate is taken (first occurrences of valid indices)
  bool take;
};
__device__ __forceinline__ ChunkPlan plan_first_chunk(const RowSource& src, long long first_chunk, int ncand) {
  const int lane = threadIdx.x & 31;
  ChunkPlan pl;
  pl.gi = lane < ncand ? first_chunk : -1;
  pl.part = lane < ncand ? find_part(src, pl.gi) : -1;
  pl.take = pl.part >= 0;
  const unsigned vmask = __ballot_sync(kFull, pl.take);
  if (pl.take) {
    const unsigned same = __match_any_sync(vmask, static_cast<unsigned long long>(pl.gi));
    pl.take = (__ffs(same) - 1) == lane;   // first occurrence inside this chunk
  }
  pl.keep = __ballot_sync(kFull, pl.take);
  return pl;
}

__device__ __forceinline__ uint32_t prefetch_group(EmbStage* st, int grp, const RowSource& src,
                                                   const long long* cand, int ncand, int cap, float* rows_grp,
                                                   int d, uint64_t* bar, const ChunkPlan& first) {
  const int lane = threadIdx.x & 31;
  const uint32_t row_bytes = static_cast<uint32_t>(d) * 4u;
  int n = 0, end = 0;
  for (int c0 = 0; c0 < ncand && n < cap; c0 += 32) {
    const int c = c0 + lane;
    long long gi = -1;
    int part = -1;
    bool take;
    unsigned keep;
    if (c0 == 0) {
      gi = first.gi;
      part = first.part;
      take = first.take;
      keep = first.keep;
    } else {
      if (c < ncand) {
        gi = cand[c];
        part = find_part(src, gi);
      }
      take = part >= 0;
      if (take)
        for (int j = 0; j < n; ++j) take &= (st->pf_idx[grp][j] != gi);
      const unsigned vmask = __ballot_sync(kFull, take);
      if (take) {
        const unsigned same = __match_any_sync(vmask, static_cast<unsigned long long>(gi));
        take = (__ffs(same) - 1) == lane;   // first occurrence inside this chunk
      }
      keep = __ballot_sync(kFull, take);
    }
    const int slot = n + __popc(keep & ((1u << lane) - 1u));
    if (take && slot < cap) {
      st->pf_idx[grp][slot] = gi;
      bulk_g2s(rows_grp + static_cast<size_t>(slot) * d, src.f32[part] + (gi - src.off[part]) * d, row_bytes, bar);
    }
    const int taken = min(cap - n, __popc(keep));
    if (taken > 0) {
      // candidate position after the last one taken in this chunk
      unsigned k2 = keep;
      for (int i = 1; i < taken; ++i) k2 &= k2 - 1;   // drop the lowest taken-1 bits
      end = c0 + __ffs(k2);
    }
    n += taken;
    __syncwarp();
  }
  if (n < cap) end = ncand;   // list exhausted
  if (lane == 0) {
    st->pf_n[grp] = n;
    st->pf_end[grp] = end;
  }
  return static_cast<uint32_t>(n) * row_bytes;
}

__global__ void __launch_bounds__(kEmbThreads, 1)
consistency_emb_pipe_kernel(const tvc_detector_params p, long long nq, int d, const ConsistencyEmbArgs a,
                            int n_stages, int stage_rows) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) uint64_t s_full[kEmbMaxStages], s_done[kEmbMaxStages], s_empty[kEmbMaxStages];
  __shared__ double s_rcp[kRcpN];
  __shared__ EmbTask s_tasks[kMaxTasks];
  __shared__ int s_ntasks;
  __shared__ unsigned s_next;
  __shared__ unsigned s_released[kEmbMaxStages];   // queries finished per stage (see the consumer loop)
  __shared__ EmbStage s_stage[kEmbMaxStages];
  __shared__ EmbLists s_lists[kEmbConsumers];
  float* s_rows = reinterpret_cast<float*>(s_raw);
  const size_t stage_floats = static_cast<size_t>(stage_rows) * d;

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = a.var ? p.n_variants : 0;
  const bool has_ret = a.ret_idx != nullptr && a.ret.nparts > 0;
  const bool gen_direct = a.gen != nullptr;
  const bool gen_idx = !gen_direct && a.gen_idx != nullptr && a.genr.nparts > 0;
  const int R = has_ret ? p.n_retrieval : 0;
  const int G = (gen_direct || gen_idx) ? p.n_generative : 0;
  const int row_var = 2, row_ret = 2 + V, row_gen = 2 + V + R;

  fill_rcp_table(s_rcp, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) {
    // task table (the same for every query): image group, variant pairs, reference pairs
    int nt = 0;
    auto add = [&](int a_row, int b_row0, int nb, int norms, int ga, int gb, int ia, int ib0, int out) {
      EmbTask t;
      t.a_row = static_cast<short>(a_row); t.b_row0 = static_cast<short>(b_row0);
      t.nb = static_cast<unsigned char>(nb); t.norms = static_cast<unsigned char>(norms);
      t.ga = static_cast<unsigned char>(ga); t.gb = static_cast<unsigned char>(gb);
      t.ia = static_cast<short>(ia); t.ib0 = static_cast<short>(ib0);
      t.out = static_cast<short>(out); t.pad = 0;
      s_tasks[nt++] = t;
    };
    // image (row 0) against every other row; d_img[row]
    for (int b = 1; b < 2 + V; b += kJB) add(0, b, min(kJB, 2 + V - b), 1, 0, 0, 0, b, b);
    for (int b = 0; b < R; b += kJB) add(0, row_ret + b, min(kJB, R - b), 1, 0, 1, 0, b, row_ret + b);
    for (int b = 0; b < G; b += kJB)
      add(0, row_gen + b, min(kJB, G - b), 1, 0, gen_direct ? 3 : 2, 0, b, row_gen + b);
    // pair tables of a group in 5 x 5 blocks (diagonal blocks: pairs i < j only)
    static_assert(TVC_MAX_REFS == TVC_MAX_VARIANTS, "one pair-table stride");
    auto add_pairs = [&](int row0, int n, int grp) {
      for (int bi = 0; bi < n; bi += kJB)
        for (int bj = bi; bj < n; bj += kJB) {
          const int na = min(kJB, n - bi), nb = min(kJB, n - bj);
          if (bi == bj && na < 2) continue;
          add(row0 + bi, row0 + bj, nb, bi == bj ? 2 : 0, grp, grp, bi, bj, 0);
          s_tasks[nt - 1].pad = static_cast<short>(na);
        }
    };
    add_pairs(row_var, V, 0);
    if (p.dedup_threshold > -1.0f) {
      add_pairs(row_ret, R, 1);
      if (gen_idx) add_pairs(row_gen, G, 2);
    }
    s_ntasks = nt;
    s_next = 0u;
    for (int s = 0; s < kEmbMaxStages; ++s) s_released[s] = 0u;
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_done[s], static_cast<uint32_t>(nt));   // one arrival per task
      mbar_init(&s_empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int ntasks = s_ntasks;
  const long long my_n = nq > blockIdx.x ? (nq - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (w >= kEmbConsumers) {
    // ===== producer warps: a query costs its producer two dependent global loads of latency, so
    // every stage has its own producer (which also keeps each stage's barrier phases in order)
    const uint32_t row_bytes = static_cast<uint32_t>(d) * 4u;
    if (w - kEmbConsumers >= n_stages) return;
    const int s = w - kEmbConsumers;     // this producer's stage
    uint32_t round = 0;
    // the first 32 candidates of each list are read one query ahead, so the dependent global load is
    // not on the stage's critical path (it completes while the producer is parked on `empty`)
    long long pre_ret = -1, pre_gen = -1;
    auto preload = [&](long long qn) {
      if (has_ret && lane < a.n_ret_cand)
        pre_ret = (reinterpret_cast<const long long*>(a.ret_idx) + qn * a.n_ret_cand)[lane];
      if (gen_idx && lane < a.n_gen_cand)
        pre_gen = (reinterpret_cast<const long long*>(a.gen_idx) + qn * a.n_gen_cand)[lane];
    };
    if (s < my_n) preload(blockIdx.x + s * static_cast<long long>(gridDim.x));
    for (long long i = s; i < my_n; i += n_stages, ++round) {
      const long long cur_ret = pre_ret, cur_gen = pre_gen;
      if (i + n_stages < my_n) preload(blockIdx.x + (i + n_stages) * static_cast<long long>(gridDim.x));
      ChunkPlan plan_ret{}, plan_gen{};
      if (has_ret) plan_ret = plan_first_chunk(a.ret, cur_ret, a.n_ret_cand);
      if (gen_idx) plan_gen = plan_first_chunk(a.genr, cur_gen, a.n_gen_cand);
      if (round > 0) mbar_wait_parked(&s_empty[s], (round - 1) & 1u);
      const long long q = blockIdx.x + i * static_cast<long long>(gridDim.x);
      EmbStage* st = &s_stage[s];
      float* rows = s_rows + static_cast<size_t>(s) * stage_floats;
      uint64_t* bar = &s_full[s];
      if (lane == 0) TVC_TRACE(i, 0);
      // rows addressed directly by q: image, text, the V variants (contiguous) and the direct
      // generative rows (contiguous): four bulk copies
      const int ndirect = 2 + V + (gen_direct ? G : 0);
      if (lane == 0) bulk_g2s(rows, a.img + q * d, row_bytes, bar);
      if (lane == 1) bulk_g2s(rows + d, a.txt + q * d, row_bytes, bar);
      if (lane == 2 && V > 0)
        bulk_g2s(rows + static_cast<size_t>(row_var) * d, a.var + q * V * d, row_bytes * V, bar);
      if (lane == 3 && gen_direct && G > 0)
        bulk_g2s(rows + static_cast<size_t>(row_gen) * d, a.gen + q * G * d, row_bytes * G, bar);
      uint32_t bytes = static_cast<uint32_t>(ndirect) * row_bytes;
      if (has_ret)
        bytes += prefetch_group(st, 0, a.ret, reinterpret_cast<const long long*>(a.ret_idx) + q * a.n_ret_cand,
                                a.n_ret_cand, R, rows + static_cast<size_t>(row_ret) * d, d, bar, plan_ret);
      if (gen_idx)
        bytes += prefetch_group(st, 1, a.genr, reinterpret_cast<const long long*>(a.gen_idx) + q * a.n_gen_cand,
                                a.n_gen_cand, G, rows + static_cast<size_t>(row_gen) * d, d, bar, plan_gen);
      if (lane == 0) {
        st->q = q;
        if (!has_ret) { st->pf_n[0] = 0; st->pf_end[0] = 0; }
        if (!gen_idx) { st->pf_n[1] = 0; st->pf_end[1] = 0; }
        st->n_gen_direct = gen_direct ? (a.g_cnt ? max(0, min(G, a.g_cnt[q])) : G) : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
      if (lane == 0) TVC_TRACE(i, 1);
    }
    return;
  }

  // ===== consumer warps: one CTA-wide queue of (query, task) units in query order
  EmbLists* L = &s_lists[w];
  const unsigned total_units = static_cast<unsigned>(my_n) * static_cast<unsigned>(ntasks);   // < 2^31 (launcher)
  while (true) {
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(&s_next, 1u);
    g = __shfl_sync(kFull, g, 0);
    if (g >= total_units) break;
    const unsigned iu = g / static_cast<unsigned>(ntasks);      // 32-bit divisions only
    const int t = static_cast<int>(g - iu * static_cast<unsigned>(ntasks));
    const unsigned round = iu / static_cast<unsigned>(n_stages);
    const int s = static_cast<int>(iu - round * static_cast<unsigned>(n_stages));
    const uint32_t par = round & 1u;
    const long long i = iu;
    EmbStage* st = &s_stage[s];
    float* rows = s_rows + static_cast<size_t>(s) * stage_floats;
    // An mbarrier wait names its phase by one parity bit, so a waiter may be at most one phase ahead.
    // With fewer tasks per query than consumer warps the unit queue can run two rounds ahead of a stage
    // that a slow finisher (the de-duplication slow path) still holds: a unit of round r would then see
    // the completed phase r - 2 and start on the previous query's rows.  Wait until every earlier round
    // of the stage has been released before naming the phase.
    if (round > 0) {
      unsigned rel = 0;
      do {
        if (lane == 0) rel = *reinterpret_cast<volatile unsigned*>(&s_released[s]);
        rel = __shfl_sync(kFull, rel, 0);
        if (rel < round) __nanosleep(200);
      } while (rel < round);
    }
    mbar_wait_parked(&s_full[s], par);
    if (lane == 0) TVC_TRACE(i, 2);
    const unsigned long long t_task0 = a.trace != nullptr ? gtime_ns() : 0ull;
    {
      const EmbTask tk = s_tasks[t];
      const int valid1 = st->pf_n[0], valid2 = st->pf_n[1], valid3 = st->n_gen_direct;
      const int va = tk.ga == 0 ? 1 << 20 : (tk.ga == 1 ? valid1 : valid2);
      const int vb = tk.gb == 0 ? 1 << 20 : (tk.gb == 1 ? valid1 : (tk.gb == 2 ? valid2 : valid3));
      const int nb = min(static_cast<int>(tk.nb), vb - tk.ib0);
      if (tk.ia < va && nb > 0) {
        if (tk.norms == 1) {
          run_task<true>(tk, nb, rows, d, st, st->d_img);
        } else {
          const int na = min(static_cast<int>(tk.pad), va - tk.ia);
          float* tab = tk.ga == 0 ? st->d_var : st->d_ref[tk.ga - 1];
          if (tk.norms == 2)
            run_block_task<true>(tk, na, nb, rows, d, tab);
          else
            run_block_task<false>(tk, na, nb, rows, d, tab);
        }
      }
    }
    __syncwarp();
    if (a.trace != nullptr && blockIdx.x == 0 && i < 512 && lane == 0)   // slot 6: longest task (ns << 8 | task)
      atomicMax(a.trace + i * 8 + 6, ((gtime_ns() - t_task0) << 8) | static_cast<unsigned long long>(t));
    if (lane == 0) mbar_arrive(&s_done[s]);     // one arrival per unit, executed or skipped
    if (t != ntasks - 1) continue;

    // ===== the warp that drew a query's last unit finishes it (the others keep drawing units)
    mbar_wait_parked(&s_done[s], par);
    if (lane == 0) TVC_TRACE(i, 3);
    const long long q = st->q;
    const int nx = V * (V - 1) / 2;
    const float thr = p.dedup_threshold;
    // Can the greedy de-duplication drop anything?  cos(i, j) > thr needs dot > 0 and
    // dot^2 > thr^2 |i|^2 |j|^2, which costs no division or square root; the test below is that
    // inequality loosened by 1e-4 (and always true for thr <= 0 or degenerate rows), one lane per row.
    bool maybe_dup = false;
    if (thr > -1.0f) {
      const float t2 = thr > 0.f ? thr * thr * (1.0f - 1e-4f) : -1.f;
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        if (grp == 0 ? !has_ret : !gen_idx) continue;
        const int n = st->pf_n[grp], base = grp == 0 ? row_ret : row_gen;
        if (lane < n) {
          const float ns = st->nrm2[base + lane];
          for (int j = 0; j < lane; ++j) {
            const float dot = st->d_ref[grp][j * TVC_MAX_REFS + lane], nn = st->nrm2[base + j] * ns;
            maybe_dup |= t2 < 0.f || nn < 1e-12f || (dot > 0.f && dot * dot > t2 * nn);
          }
        }
      }
    }
    maybe_dup = __any_sync(kFull, maybe_dup);
    float s0;
    int nr = 0, ng = 0;
    if (!maybe_dup) {
      // fast path (the common case): nothing can be dropped, so the first min(cap, prefetched) rows are
      // the references.  Copy the ~50 scalars still needed, hand the stage back, then do the divisions.
      nr = has_ret ? min(R, st->pf_n[0]) : 0;
      ng = gen_direct ? st->n_gen_direct : (gen_idx ? min(G, st->pf_n[1]) : 0);
      for (int r = lane; r < stage_rows; r += 32) {
        L->sq[r] = st->nrm2[r];
        L->dimg[r] = st->d_img[r];
      }
      for (int e = lane; e < nx; e += 32) {
        int ii = 0, rem = e;     // e-th pair (ii < jj) in row-major order of the upper triangle
        while (rem >= V - 1 - ii) {
          rem -= V - 1 - ii;
          ++ii;
        }
        L->sx[e] = st->d_var[ii * TVC_MAX_VARIANTS + ii + 1 + rem];
      }
      __syncwarp();
      if (lane == 0) {
        atomicAdd(&s_released[s], 1u);
        mbar_arrive(&s_empty[s]);
      }
      if (lane == 0) TVC_TRACE(i, 4);
      for (int r = lane; r < stage_rows; r += 32) L->sq[r] = sqrtf(L->sq[r]);
      __syncwarp();
      const float sq_img = L->sq[0];
      s0 = L->dimg[1] / fmaxf(sq_img * L->sq[1], 1e-8f);
      if (lane < V) L->sv[lane] = L->dimg[row_var + lane] / fmaxf(sq_img * L->sq[row_var + lane], 1e-8f);
      for (int e = lane; e < nx; e += 32) {
        int ii = 0, rem = e;
        while (rem >= V - 1 - ii) {
          rem -= V - 1 - ii;
          ++ii;
        }
        L->sx[e] = L->sx[e] / fmaxf(L->sq[row_var + ii] * L->sq[row_var + ii + 1 + rem], 1e-8f);
      }
      if (lane < nr) L->sr[lane] = L->dimg[row_ret + lane] / fmaxf(sq_img * L->sq[row_ret + lane], 1e-8f);
      if (lane < ng) L->sg[lane] = L->dimg[row_gen + lane] / fmaxf(sq_img * L->sq[row_gen + lane], 1e-8f);
      __syncwarp();
    } else {
      for (int r = lane; r < stage_rows; r += 32) L->sq[r] = sqrtf(st->nrm2[r]);
      __syncwarp();
      const float sq_img = L->sq[0];
      s0 = st->d_img[1] / fmaxf(sq_img * L->sq[1], 1e-8f);
      if (lane < V) L->sv[lane] = st->d_img[row_var + lane] / fmaxf(sq_img * L->sq[row_var + lane], 1e-8f);
      for (int e = lane; e < nx; e += 32) {
        int ii = 0, rem = e;
        while (rem >= V - 1 - ii) {
          rem -= V - 1 - ii;
          ++ii;
        }
        const int jj = ii + 1 + rem;
        L->sx[e] = st->d_var[ii * TVC_MAX_VARIANTS + jj] / fmaxf(L->sq[row_var + ii] * L->sq[row_var + jj], 1e-8f);
      }
      if (has_ret)
        nr = select_from_stage(p, st, 0, rows, d, row_ret, a.ret,
                               reinterpret_cast<const long long*>(a.ret_idx) + q * a.n_ret_cand, a.n_ret_cand, R,
                               L->sr, L);
      if (gen_direct) {
        ng = st->n_gen_direct;
        if (lane < ng) L->sg[lane] = st->d_img[row_gen + lane] / fmaxf(sq_img * L->sq[row_gen + lane], 1e-8f);
      } else if (gen_idx) {
        ng = select_from_stage(p, st, 1, rows, d, row_gen, a.genr,
                               reinterpret_cast<const long long*>(a.gen_idx) + q * a.n_gen_cand, a.n_gen_cand, G,
                               L->sg, L);
      }
      // everything still needed lives in this warp's lists: hand the stage back
      // (the slow path may have written rows with ordinary stores; the next writer is a bulk copy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        atomicAdd(&s_released[s], 1u);
        mbar_arrive(&s_empty[s]);
      }
      if (lane == 0) TVC_TRACE(i, 4);
    }
    // similarity lists -> global; the statistics kernel (consistency_sims_kernel) takes it from there
    const int Vp = p.n_variants, Rp = p.n_retrieval, Gp = p.n_generative;
    if (lane == 0) {
      a.w_s0[q] = s0;
      a.w_rcnt[q] = nr;
      a.w_gcnt[q] = ng;
    }
    if (lane < Vp) {
      const float x = lane < V ? L->sv[lane] : 0.f;
      a.w_sv[q * Vp + lane] = x;
      if (a.out_sv) a.out_sv[q * Vp + lane] = x;
    }
    if (lane < Rp) {
      const float x = lane < nr ? L->sr[lane] : 0.f;
      a.w_sr[q * Rp + lane] = x;
      if (a.out_sr) a.out_sr[q * Rp + lane] = x;
    }
    if (lane < Gp) {
      const float x = lane < ng ? L->sg[lane] : 0.f;
      a.w_sg[q * Gp + lane] = x;
      if (a.out_sg) a.out_sg[q * Gp + lane] = x;
    }
    for (int e = lane; e < nx; e += 32) a.w_sx[q * nx + e] = L->sx[e];
    __syncwarp();
    if (lane == 0) TVC_TRACE(i, 5);
  }
}

// ------------------------------------------------------------------------------- README reference-vector rule
// README.md:474-482, 846: per text variant v the top-k retrieved gallery rows and the m generated rows
// are averaged into a per-variant reference vector r_v, the variants' vectors are averaged into the
// Reference Vector r, S_v = cos(image, r_v), sigma = std_v(S_v) and the sample is adversarial iff
// sigma > threshold (confidence = sigma).  Every gathered row is used exactly once, so nothing is
// staged: one block per query, one warp per variant streams its k + m rows with 128-bit loads into a
// register accumulator (two rows in flight), dots it with the image row and leaves r_v in shared
// memory for the cross-variant mean.  Bytes per query: 4 d (1 + V (k + m)).
constexpr int kRvSeg = 1024;            // floats of a row handled per pass (8 float4 per lane)
__global__ void __launch_bounds__(TVC_MAX_VARIANTS * 32)
reference_vector_kernel(long long nq, int d, int V, const float* __restrict__ img, const RowSource src,
                        const long long* __restrict__ ret_idx, int k, const float* __restrict__ gen, int m,
                        float sigma_threshold, float* __restrict__ out_s, float* __restrict__ out_ref,
                        float* __restrict__ out_sigma, uint8_t* __restrict__ flags) {
  extern __shared__ __align__(16) float s_rv[];        // [V][seg] per-variant mean segment
  __shared__ float s_part[TVC_MAX_VARIANTS][3];         // img.sum, |sum|^2, (unused)
  __shared__ float s_ref[3];                            // img.r, |r|^2, |img|^2
  __shared__ int s_cnt[TVC_MAX_VARIANTS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (d & 3) == 0;
  for (long long q = blockIdx.x; q < nq; q += gridDim.x) {
    float dot_v = 0.f, nrm_v = 0.f;       // this warp's variant: img . sum_v and |sum_v|^2
    float dot_r = 0.f, nrm_r = 0.f, nrm_i = 0.f;   // warp 0: img . r, |r|^2, |img|^2
    // resolve this variant's rows once (lane j holds row j's pointer; k + m <= 32 is enforced by the host)
    const float* my_row = nullptr;
    if (lane < k) {
      const long long gi = ret_idx[(q * V + w) * k + lane];
      const int part = find_part(src, gi);
      if (part >= 0 && src.f32[part] != nullptr) my_row = src.f32[part] + (gi - src.off[part]) * d;
    } else if (lane < k + m) {
      my_row = gen + ((q * V + w) * m + (lane - k)) * d;
    }
    const unsigned have = __ballot_sync(kFull, my_row != nullptr);
    const int cnt = __popc(have);
    if (lane == 0) s_cnt[w] = cnt;
    const float inv_cnt = cnt > 0 ? 1.0f / static_cast<float>(cnt) : 0.f;
    const float* irow = img + q * d;
    for (int seg0 = 0; seg0 < d; seg0 += kRvSeg) {
      const int seg = min(kRvSeg, d - seg0);
      float4 acc[kRvSeg / 128];
#pragma unroll
      for (int c = 0; c < kRvSeg / 128; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned todo = have;
      while (todo) {
        const int j0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const int j1 = todo ? __ffs(todo) - 1 : -1;
        if (j1 >= 0) todo &= todo - 1;
        const float* r0 = reinterpret_cast<const float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), j0)) + seg0;
        const float* r1 = j1 >= 0 ? reinterpret_cast<const float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), j1)) + seg0 : nullptr;
        if (vec) {
          float4 x[kRvSeg / 128], y[kRvSeg / 128];
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            const int e = (c * 32 + lane) * 4;
            x[c] = e < seg ? __ldg(reinterpret_cast<const float4*>(r0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            y[c] = (r1 && e < seg) ? __ldg(reinterpret_cast<const float4*>(r1 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            acc[c].x += x[c].x; acc[c].y += x[c].y; acc[c].z += x[c].z; acc[c].w += x[c].w;
            acc[c].x += y[c].x; acc[c].y += y[c].y; acc[c].z += y[c].z; acc[c].w += y[c].w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            float* ac = reinterpret_cast<float*>(&acc[c]);
            for (int t = 0; t < 4; ++t) {
              const int e = (c * 32 + lane) * 4 + t;
              if (e < seg) ac[t] += r0[e] + (r1 ? r1[e] : 0.f);
            }
          }
        }
      }
      // dot with the image segment, norm of the sum, and (when the Reference Vector is wanted) the
      // per-variant MEAN segment to shared memory
      const bool want_ref = out_ref != nullptr;
#pragma unroll
      for (int c = 0; c < kRvSeg / 128; ++c) {
        const int e0 = (c * 32 + lane) * 4;
        if (e0 >= seg) continue;
        float xi[4];
        if (vec) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(irow + seg0 + e0));
          xi[0] = t4.x; xi[1] = t4.y; xi[2] = t4.z; xi[3] = t4.w;
        } else {
          for (int t = 0; t < 4; ++t) xi[t] = e0 + t < seg ? irow[seg0 + e0 + t] : 0.f;
        }
        const float* ac = reinterpret_cast<const float*>(&acc[c]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (e0 + t < seg) {
            dot_v = fmaf(xi[t], ac[t], dot_v);
            nrm_v = fmaf(ac[t], ac[t], nrm_v);
            if (!want_ref) nrm_i = fmaf(xi[t], xi[t], nrm_i);
            if (want_ref) s_rv[w * kRvSeg + e0 + t] = ac[t] * inv_cnt;
          }
        }
      }
      if (!want_ref) continue;
      __syncthreads();
      if (w == 0) {
        for (int e = lane; e < seg; e += 32) {
          float r = 0.f;
          int nv = 0;
          for (int v = 0; v < V; ++v)
            if (s_cnt[v] > 0) {
              r += s_rv[v * kRvSeg + e];
              ++nv;
            }
          r = nv > 0 ? r / static_cast<float>(nv) : 0.f;
          const float xi = irow[seg0 + e];
          dot_r = fmaf(xi, r, dot_r);
          nrm_r = fmaf(r, r, nrm_r);
          nrm_i = fmaf(xi, xi, nrm_i);
        }
      }
      __syncthreads();
    }
    dot_v = warp_sum(dot_v);
    nrm_v = warp_sum(nrm_v);
    if (w == 0) {
      dot_r = warp_sum(dot_r);
      nrm_r = warp_sum(nrm_r);
      nrm_i = warp_sum(nrm_i);     // (without the Reference Vector every warp accumulated |img|^2 itself)
      if (lane == 0) {
        s_ref[0] = dot_r;
        s_ref[1] = nrm_r;
        s_ref[2] = nrm_i;
      }
    }
    if (lane == 0) {
      s_part[w][0] = dot_v;
      s_part[w][1] = nrm_v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float ni = s_ref[2];
      double sum = 0.0;
      int n = 0;
      float sv[TVC_MAX_VARIANTS];
      for (int v = 0; v < V; ++v) {
        sv[v] = 0.f;
        if (s_cnt[v] > 0) {
          sv[v] = cos_of(s_part[v][0], ni, s_part[v][1]);
          sum += sv[v];
          ++n;
        }
        out_s[q * V + v] = sv[v];
      }
      double sigma = 0.0;
      if (n > 0) {
        const double mu = sum / n;
        double acc2 = 0.0;
        for (int v = 0; v < V; ++v)
          if (s_cnt[v] > 0) acc2 = fma(sv[v] - mu, sv[v] - mu, acc2);
        sigma = sqrt(acc2 / n);
      }
      out_sigma[q] = static_cast<float>(sigma);
      if (out_ref) out_ref[q] = n > 0 ? cos_of(s_ref[0], ni, s_ref[1]) : 0.f;
      flags[q] = sigma > static_cast<double>(sigma_threshold) ? TVC_FLAG_SIGMA_ADV : 0;
    }
    __syncthreads();
  }
}

// Same rule when the Reference Vector itself is not asked for: the (query, variant) pairs are independent,
// so every warp takes pairs from a flat index space (no block-level synchronisation at all) and a
// second, tiny kernel turns S [Q, V] into sigma and the decision.
__global__ void __launch_bounds__(256, 4)
reference_vector_flat_kernel(long long nq, int d, int V, const float* __restrict__ img, const RowSource src,
                             const long long* __restrict__ ret_idx, int k, const float* __restrict__ gen, int m,
                             float* __restrict__ out_s, uint8_t* __restrict__ valid) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (d & 3) == 0;
  for (long long item = warp0; item < nq * V; item += nwarps) {
    const long long q = item / V;
    const float* my_row = nullptr;
    if (lane < k) {
      const long long gi = ret_idx[item * k + lane];
      const int part = find_part(src, gi);
      if (part >= 0 && src.f32[part] != nullptr) my_row = src.f32[part] + (gi - src.off[part]) * d;
    } else if (lane < k + m) {
      my_row = gen + (item * m + (lane - k)) * d;
    }
    const unsigned have = __ballot_sync(kFull, my_row != nullptr);
    const float* irow = img + q * d;
    float dot_v = 0.f, nrm_v = 0.f, nrm_i = 0.f;
    for (int seg0 = 0; seg0 < d; seg0 += kRvSeg) {
      const int seg = min(kRvSeg, d - seg0);
      float4 acc[kRvSeg / 128];
#pragma unroll
      for (int c = 0; c < kRvSeg / 128; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned todo = have;
      while (todo) {
        const int j0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const int j1 = todo ? __ffs(todo) - 1 : -1;
        if (j1 >= 0) todo &= todo - 1;
        const float* r0 = reinterpret_cast<const float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), j0)) + seg0;
        const float* r1 = j1 >= 0 ? reinterpret_cast<const float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), j1)) + seg0 : nullptr;
        if (vec) {
          float4 x[kRvSeg / 128], y[kRvSeg / 128];
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            const int e = (c * 32 + lane) * 4;
            x[c] = e < seg ? __ldcs(reinterpret_cast<const float4*>(r0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            y[c] = (r1 && e < seg) ? __ldcs(reinterpret_cast<const float4*>(r1 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            acc[c].x += x[c].x; acc[c].y += x[c].y; acc[c].z += x[c].z; acc[c].w += x[c].w;
            acc[c].x += y[c].x; acc[c].y += y[c].y; acc[c].z += y[c].z; acc[c].w += y[c].w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < kRvSeg / 128; ++c) {
            float* ac = reinterpret_cast<float*>(&acc[c]);
            for (int t = 0; t < 4; ++t) {
              const int e = (c * 32 + lane) * 4 + t;
              if (e < seg) ac[t] += r0[e] + (r1 ? r1[e] : 0.f);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kRvSeg / 128; ++c) {
        const int e0 = (c * 32 + lane) * 4;
        if (e0 >= seg) continue;
        float xi[4];
        if (vec) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(irow + seg0 + e0));
          xi[0] = t4.x; xi[1] = t4.y; xi[2] = t4.z; xi[3] = t4.w;
        } else {
          for (int t = 0; t < 4; ++t) xi[t] = e0 + t < seg ? irow[seg0 + e0 + t] : 0.f;
        }
        const float* ac = reinterpret_cast<const float*>(&acc[c]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (e0 + t < seg) {
            dot_v = fmaf(xi[t], ac[t], dot_v);
            nrm_v = fmaf(ac[t], ac[t], nrm_v);
            nrm_i = fmaf(xi[t], xi[t], nrm_i);
          }
      }
    }
    dot_v = warp_sum(dot_v);
    nrm_v = warp_sum(nrm_v);
    nrm_i = warp_sum(nrm_i);
    if (lane == 0) {
      out_s[item] = have ? cos_of(dot_v, nrm_i, nrm_v) : 0.f;
      valid[item] = have ? 1 : 0;
    }
  }
}

__global__ void reference_sigma_kernel(long long nq, int V, const float* __restrict__ s,
                                       const uint8_t* __restrict__ valid, float sigma_threshold,
                                       float* __restrict__ out_sigma, uint8_t* __restrict__ flags) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double sum = 0.0;
  int n = 0;
  for (int v = 0; v < V; ++v)
    if (valid[q * V + v]) {
      sum += s[q * V + v];
      ++n;
    }
  double sigma = 0.0;
  if (n > 0) {
    const double mu = sum / n;
    double acc2 = 0.0;
    for (int v = 0; v < V; ++v)
      if (valid[q * V + v]) acc2 = fma(s[q * V + v] - mu, s[q * V + v] - mu, acc2);
    sigma = sqrt(acc2 / n);
  }
  out_sigma[q] = static_cast<float>(sigma);
  flags[q] = sigma > static_cast<double>(sigma_threshold) ? TVC_FLAG_SIGMA_ADV : 0;
}

}  // namespace

// =============================================================================== launchers
cudaError_t launch_consistency_sims(const tvc_detector_params& p, int64_t q, const float* s0,
                                    const float* sv, const float* sr, const int32_t* r_cnt,
                                    const float* sg, const int32_t* g_cnt, const float* sxv,
                                    float* scores, uint8_t* flags, cudaStream_t stream) {
  if (q <= 0) return cudaSuccess;
  const int V = p.n_variants, R = p.n_retrieval, G = p.n_generative;
  const int X = sxv ? V * (V - 1) / 2 : 0;
  auto up4 = [](int x) { return (x + 3) & ~3; };
  const size_t floats = up4(kSimsBlock * V) + up4(kSimsBlock * R) + up4(kSimsBlock * G) +
                        up4(kSimsBlock * X) + 16 + kOutStride * kSimsBlock;
  const size_t smem = floats * 4;
  static const int minb = [] {
    const char* e = getenv("TVC_SIMS_MIN_BLOCKS");
    const int v = e ? atoi(e) : 0;
    return (v == 4 || v == 6 || v == 8) ? v : kSimsMinBlocks;
  }();
  const int grid = static_cast<int>((q + kSimsBlock - 1) / kSimsBlock);
  auto go = [&](auto kernel, SmemAttrOnce& once) -> cudaError_t {
    if (cudaError_t e = once.ensure(reinterpret_cast<const void*>(kernel), 160 * 1024); e != cudaSuccess) return e;
    kernel<<<grid, kSimsBlock, smem, stream>>>(p, q, s0, sv, sr, r_cnt, sg, g_cnt, sxv, scores, flags);
    return cudaSuccess;
  };
  static SmemAttrOnce c4, c6, c8;   // per (kernel, device): a second context on another GPU sets its own
  cudaError_t le = minb == 4 ? go(consistency_sims_kernel<4>, c4)
                   : (minb == 6 ? go(consistency_sims_kernel<6>, c6) : go(consistency_sims_kernel<8>, c8));
  if (le != cudaSuccess) return le;
  note_launch();
  return cudaGetLastError();
}

static bool rows_f32_aligned(const RowSource& s) {
  for (int i = 0; i < s.nparts; ++i)
    if (s.f32[i] == nullptr || (reinterpret_cast<uintptr_t>(s.f32[i]) & 15u) != 0) return false;
  return true;
}

cudaError_t launch_consistency_emb(const tvc_detector_params& p, int64_t q, int d,
                                   const ConsistencyEmbArgs& a, float* scores, uint8_t* flags,
                                   int sm_count, int force_generic, cudaStream_t stream) {
  if (q <= 0) return cudaSuccess;
  // ---- pipelined kernel: fp32 rows, 16-byte aligned, at least two stages of rows in shared memory
  const int V = a.var ? p.n_variants : 0;
  const bool has_ret = a.ret_idx != nullptr && a.ret.nparts > 0;
  const bool gen_direct = a.gen != nullptr;
  const bool gen_idx = !gen_direct && a.gen_idx != nullptr && a.genr.nparts > 0;
  const int R = has_ret ? p.n_retrieval : 0;
  const int G = (gen_direct || gen_idx) ? p.n_generative : 0;
  const int stage_rows = 2 + V + R + G;
  const size_t stage_bytes = static_cast<size_t>(stage_rows) * d * 4;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  bool pipe = !force_generic && (d % 4) == 0 && al16(a.img) && al16(a.txt) && (!a.var || al16(a.var)) &&
              (!a.gen || al16(a.gen)) && (!has_ret || rows_f32_aligned(a.ret)) &&
              (!gen_idx || rows_f32_aligned(a.genr));
  // dynamic shared memory left for the stage ring after the kernel's static tables
  static std::atomic<size_t> static_smem{0};   // the kernel's static tables: a property of the code, not of the device
  if (static_smem.load(std::memory_order_acquire) == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, consistency_emb_pipe_kernel);
    if (e != cudaSuccess) return e;
    static_smem.store(fa.sharedSizeBytes + 1, std::memory_order_release);
  }
  const size_t smem_budget = 227 * 1024 - (static_smem.load(std::memory_order_acquire) - 1) - 1024;
  static SmemAttrOnce configured_p;            // the opt-in itself is per (kernel, device)
  if (cudaError_t e = configured_p.ensure(reinterpret_cast<const void*>(consistency_emb_pipe_kernel),
                                          static_cast<int>(smem_budget)); e != cudaSuccess)
    return e;
  int n_stages = static_cast<int>(smem_budget / stage_bytes);
  if (n_stages > kEmbMaxStages) n_stages = kEmbMaxStages;
  if (n_stages < 2) pipe = false;
  if (q / (sm_count > 0 ? sm_count : 1) + 1 >= (1ll << 31) / kMaxTasks) pipe = false;   // 32-bit unit counter
  if (pipe) {
    long long blocks = q < sm_count ? q : sm_count;
    consistency_emb_pipe_kernel<<<static_cast<int>(blocks), kEmbThreads, n_stages * stage_bytes, stream>>>(
        p, q, d, a, n_stages, stage_rows);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // statistics + decisions from the similarity lists the kernel left in the workspace
    return launch_consistency_sims(p, q, a.w_s0, V > 0 ? a.w_sv : nullptr, has_ret ? a.w_sr : nullptr, a.w_rcnt,
                                   G > 0 ? a.w_sg : nullptr, a.w_gcnt, V > 1 ? a.w_sx : nullptr, scores, flags,
                                   stream);
  }
  // ---- generic kernel
  int rows_cap = p.n_retrieval > p.n_generative ? p.n_retrieval : p.n_generative;
  if (rows_cap < 1) rows_cap = 1;
  const size_t row_floats = (static_cast<size_t>(1 + rows_cap) * d + 3) & ~static_cast<size_t>(3);
  const size_t per_warp = (row_floats + kListFloats + 2 * TVC_MAX_REFS) * 4;
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  if (per_warp * warps > 220 * 1024) return cudaErrorInvalidValue;
  static SmemAttrOnce configured_g;   // per (kernel, device): a second context on another GPU sets its own
  if (cudaError_t e = configured_g.ensure(reinterpret_cast<const void*>(consistency_emb_generic_kernel), 220 * 1024); e != cudaSuccess)
    return e;
  long long blocks = (q + warps - 1) / warps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  consistency_emb_generic_kernel<<<static_cast<int>(blocks), warps * 32, per_warp * warps, stream>>>(
      p, q, d, a, scores, flags, rows_cap);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_reference_vector(int64_t q, int d, int v, const float* img, const RowSource& src,
                                    const int64_t* ret_idx, int k, const float* gen, int m, float sigma_threshold,
                                    float* out_s, float* out_ref, float* out_sigma, uint8_t* flags,
                                    uint8_t* valid_ws, int sm_count, cudaStream_t stream) {
  if (q <= 0) return cudaSuccess;
  if (v < 1 || v > TVC_MAX_VARIANTS || k + m > 32 || k < 0 || m < 0) return cudaErrorInvalidValue;
  if (out_ref == nullptr && valid_ws != nullptr) {
    // independent (query, variant) pairs + a tiny sigma pass
    long long blocks = (q * v + 7) / 8;
    const long long cap = static_cast<long long>(sm_count) * 8;
    if (blocks > cap) blocks = cap;
    reference_vector_flat_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
        q, d, v, img, src, reinterpret_cast<const long long*>(ret_idx), k, gen, m, out_s, valid_ws);
    note_launch();
    reference_sigma_kernel<<<static_cast<int>((q + 255) / 256), 256, 0, stream>>>(q, v, out_s, valid_ws,
                                                                                  sigma_threshold, out_sigma, flags);
    note_launch();
    return cudaGetLastError();
  }
  const size_t smem = static_cast<size_t>(v) * kRvSeg * 4;
  static SmemAttrOnce configured;   // per (kernel, device): a second context on another GPU sets its own
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(reference_vector_kernel), TVC_MAX_VARIANTS * kRvSeg * 4); e != cudaSuccess)
    return e;
  long long blocks = static_cast<long long>(sm_count) * 8;
  if (blocks > q) blocks = q;
  reference_vector_kernel<<<static_cast<int>(blocks), v * 32, smem, stream>>>(
      q, d, v, img, src, reinterpret_cast<const long long*>(ret_idx), k, gen, m, sigma_threshold, out_s, out_ref,
      out_sigma, flags);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tvc
