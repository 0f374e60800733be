// Kernel (a): fused cosine-similarity GEMM + per-row top-k for sm_100a.
//
//   S[m, n] = sum_k Q[m, k] * G[n, k]      (bf16 operands, fp32 accumulate in TMEM)
//   per query row m: the KP largest S[m, n] over one gallery range, ordered (value desc, n asc)
//
// Replaces the reference's `index.search(q, k)` (src/retrieval.py:652-656), the sklearn cosine +
// full argsort fallback (:669-671), the ReferenceBank dot + argsort (src/ref_bank.py:475-484,
// 197-203) and the broadcast cosine + topk of compute_hubness (src/attacks/hubness_attack.py:
// 482-489).  The [m, n] similarity matrix is never written to memory.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0 (1 thread)  TMA producer: Q tile [128 x 64] + G tile [256 x 64] bf16 per k-block into a
//                      4-deep 128B-swizzled smem ring, completion on `full` mbarriers.
//   warp 1 (1 thread)  MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, four per
//                      k-block, accumulating into one of two 256-column TMEM buffers;
//                      tcgen05.commit releases the smem slot / publishes the accumulator.
//   warp 2             TMEM allocator (512 columns).
//   warps 4-7          epilogue: thread t owns TMEM lane t = query row t of the tile.  It reads the
//                      accumulator 32 columns at a time (tcgen05.ld 32x32b.x32), compares the chunk
//                      maximum with its current KP-th best and only on a hit walks the chunk and
//                      bubble-inserts into a sorted in-register list.  Runs under the MMA of the
//                      next gallery tile (double-buffered accumulator).
// Work unit = (query tile, gallery range); units are dealt round-robin so that CTAs running at the
// same time stream the same gallery range (each G tile is fetched from HBM once per wave and hit
// in L2 by the others).
#include "tvc_internal.h"
#include "tvc_ptx.cuh"
#include "tvc_topk.cuh"

namespace tvc {

namespace {

constexpr int kEpiWarp0 = 4;                    // first epilogue warp
constexpr int kSmemA = 0;
constexpr int kSmemB = kSmemA + kStages * kABytes;
constexpr int kSmemStage = kSmemB + kStages * kBBytes;
constexpr int kSmemBar = kSmemStage + kStageFloats * 4;
constexpr int kSmemTotal = kSmemBar + 256 + 1024;  // + barrier block + alignment slack

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int KP>
__global__ void __launch_bounds__(kThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_g, const SearchPlan p,
                 float* __restrict__ cand_val, int32_t* __restrict__ cand_idx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem + kSmemA;
  uint8_t* sB = smem + kSmemB;
  float* sStage = reinterpret_cast<float*>(smem + kSmemStage);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kSmemBar);

  // warp index in a form the compiler can prove warp-uniform: the two issue loops run on the uniform datapath, one
  // elected lane issues (see tvc_gemm_topk_pair.cu)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total_units = plan_units(p);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const SearchUnit un = plan_unit(p, u);
      const int mt = un.mt, t0 = un.t0, t1 = un.t1;
      for (int nt = t0; nt < t1; ++nt) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars->full[stage], kABytes + kBBytes);
            tma_load_2d(&tmap_q, &bars->full[stage], sA + stage * kABytes, kb * kBK, mt * kBM,
                        kEvictLast);
            tma_load_2d(&tmap_g, &bars->full[stage], sB + stage * kBBytes, kb * kBK, nt * kBN,
                        kEvictNormal);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kBM, kBN);
    const uint32_t sA_base = smem_u32(sA), sB_base = smem_u32(sB);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const SearchUnit un = plan_unit(p, u);
      const int t0 = un.t0, t1 = un.t1;
      for (int nt = t0; nt < t1; ++nt) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128_kmajor(sA_base + static_cast<uint32_t>(stage * kABytes));
          const uint64_t db = umma_desc_sw128_kmajor(sB_base + static_cast<uint32_t>(stage * kBBytes));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance the start address by 16 bf16 = 32 bytes (>>4 = 2) inside the swizzle row
              umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                           idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&bars->empty[stage]);  // slot reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit(&bars->tmem_full[acc]);  // accumulator complete
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ top-k epilogue
    const int q4 = warp & 3;          // TMEM lane quarter this warp may read
    const int row_in_tile = q4 * 32 + lane;
    float* my_stage = sStage + row_in_tile;  // element j at my_stage[j * kBM]
    TopList<KP> top;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const SearchUnit un = plan_unit(p, u);
      const int split = un.split, mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int row = mt * kBM + row_in_tile;
      const long long self_col = p.skip_self ? static_cast<long long>(row) + p.self_offset : -1ll;
      top.reset();
      float thr = -INFINITY;
      for (int nt = t0; nt < t1; ++nt) {
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr =
            tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * kBN);
        if (!(p.debug & 1)) topk_consume_tile<KP>(top, thr, t_addr, my_stage, nt * kBN, p.n_rows, self_col, p.debug);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (row < p.m_rows) topk_store<KP>(top, cand_val, cand_idx,
                                       plan_cand_base(static_cast<long long>(p.full_tiles) * kBM, p.splits, KP, row, split));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// Same mainloop, plain store epilogue: writes the fp32 similarity tile (compute_similarity_matrix,
// src/retrieval.py:682-722; batch_cosine_similarity, src/utils/metrics.py:144-164).
__global__ void __launch_bounds__(kThreads, 1)
gemm_store_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_g, int m_rows, int n_rows, int kblocks,
                  int m_tiles, int n_tiles, float* __restrict__ out, long long ld_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem + kSmemA;
  uint8_t* sB = smem + kSmemB;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kSmemBar);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int total_units = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int nt = u / m_tiles, mt = u - nt * m_tiles;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&bars->empty[stage], phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[stage], kABytes + kBBytes);
          tma_load_2d(&tmap_q, &bars->full[stage], sA + stage * kABytes, kb * kBK, mt * kBM,
                      kEvictNormal);
          tma_load_2d(&tmap_g, &bars->full[stage], sB + stage * kBBytes, kb * kBK, nt * kBN,
                      kEvictNormal);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kBM, kBN);
    const uint32_t sA_base = smem_u32(sA), sB_base = smem_u32(sB);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128_kmajor(sA_base + static_cast<uint32_t>(stage * kABytes));
        const uint64_t db = umma_desc_sw128_kmajor(sB_base + static_cast<uint32_t>(stage * kBBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                         idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit(&bars->tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= kEpiWarp0) {
    const int q4 = warp & 3;
    const int row_in_tile = q4 * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int nt = u / m_tiles, mt = u - nt * m_tiles;
      const int row = mt * kBM + row_in_tile;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr =
          tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * kBN);
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        const int col0 = nt * kBN + c * 32;
        if (row < m_rows) {
          float* o = out + static_cast<long long>(row) * ld_out + col0;
          if (col0 + 32 <= n_rows && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(o)[j] =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                              __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < n_rows) o[j] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KP>
cudaError_t launch_kp(const CUtensorMap& tq, const CUtensorMap& tg, const SearchPlan& plan,
                      float* cv, int32_t* ci, cudaStream_t stream) {
  static SmemAttrOnce configured;   // per (kernel, device): a second context on another GPU sets its own
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(gemm_topk_kernel<KP>), kSmemTotal); e != cudaSuccess)
    return e;
  gemm_topk_kernel<KP><<<plan.grid, kThreads, kSmemTotal, stream>>>(tq, tg, plan, cv, ci);
  note_launch();
  return cudaGetLastError();
}

}  // namespace

SearchPlan make_search_plan(int64_t m, int64_t n, int d_pad, int k, int sm_count, bool pair) {
  SearchPlan p{};
  p.m_rows = static_cast<int>(m);
  p.n_rows = static_cast<int>(n);
  p.kblocks = d_pad / kBK;
  p.pair = pair ? 1 : 0;
  const int tile_m = pair ? 2 * kBM : kBM;       // a CTA pair owns 256 query rows
  const int workers = pair ? sm_count / 2 : sm_count;
  p.m_tiles = static_cast<int>((m + tile_m - 1) / tile_m);
  p.n_tiles = static_cast<int>((n + kBN - 1) / kBN);
  p.kp = k <= 10 ? 16 : (k <= 26 ? 32 : 64);
  // Whole waves of query tiles run unsplit (see SearchPlan); what is left - fewer tiles than workers - is cut
  // into S gallery ranges.  S minimises waves * (unit time) where a unit costs its tiles plus ~4 tile-times of
  // fixed work (pipeline fill/drain, candidate store, its share of the re-rank) plus a quarter tile for each of
  // the first 2*KP tiles, during which nearly every 32x32 chunk takes the epilogue's slow path.  (In
  // quarter-tile units; ties go to fewer ranges.)
  p.full_tiles = (p.m_tiles / workers) * workers;
  p.rem_tiles = p.m_tiles - p.full_tiles;
  const int max_s = p.rem_tiles == 0 ? 1 : (p.n_tiles < 1024 ? p.n_tiles : 1024);
  long long best_cost = -1;
  int best_s = 1, best_tps = p.n_tiles;
  for (int s = 1; s <= max_s; ++s) {
    const int tps = (p.n_tiles + s - 1) / s;
    const int s_eff = (p.n_tiles + tps - 1) / tps;
    if (s_eff != s) continue;
    const long long units = static_cast<long long>(p.rem_tiles) * s_eff;
    const long long waves = (units + workers - 1) / workers;
    const long long warm = tps < 2 * p.kp ? tps : 2 * p.kp;
    const long long cost = waves * (4ll * tps + 16 + warm) * 64 + s_eff;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_s = s_eff;
      best_tps = tps;
    }
  }
  p.splits = best_s;
  p.tiles_per_split = best_tps;
  const long long units = plan_units(p);
  p.grid = static_cast<int>(units < workers ? units : workers);
  if (p.grid < 1) p.grid = 1;
  if (pair) p.grid *= 2;
  return p;
}

cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g,
                             const SearchPlan& plan, float* cand_val, int32_t* cand_idx,
                             cudaStream_t stream) {
  switch (plan.kp) {
    case 16:
      return launch_kp<16>(tmap_q, tmap_g, plan, cand_val, cand_idx, stream);
    case 32:
      return launch_kp<32>(tmap_q, tmap_g, plan, cand_val, cand_idx, stream);
    case 64:
      return launch_kp<64>(tmap_q, tmap_g, plan, cand_val, cand_idx, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

cudaError_t launch_gemm_store(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g, int m, int n,
                              int kblocks, float* out, int64_t ld_out, int sm_count,
                              cudaStream_t stream) {
  static SmemAttrOnce configured;   // per (kernel, device): a second context on another GPU sets its own
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(gemm_store_kernel), kSmemTotal); e != cudaSuccess)
    return e;
  const int m_tiles = (m + kBM - 1) / kBM, n_tiles = (n + kBN - 1) / kBN;
  const long long units = static_cast<long long>(m_tiles) * n_tiles;
  const int grid = static_cast<int>(units < sm_count ? units : sm_count);
  gemm_store_kernel<<<grid, kThreads, kSmemTotal, stream>>>(tmap_q, tmap_g, m, n, kblocks, m_tiles,
                                                           n_tiles, out, ld_out);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tvc
