// Kernel (a), CTA-pair revision: the same fused GEMM + top-k as tvc_gemm_topk.cu, with two SMs of a
// TPC working on one 256 x 256 output tile (tcgen05 cta_group::2).
//
// Why: ncu on the single-CTA kernel shows the L2 -> SM path at ~74 % of its peak (15 TB/s) with the
// tensor pipe at 82 %: every SM pulls a full [256 x 64] gallery block per k-step.  In a pair each CTA
// stages only ITS half of the gallery tile (128 rows) and its own 128 query rows; the MMA unit of both
// SMs reads both halves.  L2 traffic per SM drops from 96 to 64 bytes/clk and the freed shared memory
// deepens the ring from 4 to 6 stages.
//
// Roles per CTA (256 threads): warp 0 one thread = TMA producer (own A half + own B half, completion
// bytes signalled on the LEADER's full barrier); warp 1 one thread = MMA issuer (leader CTA only,
// M = 256, N = 256, K = 16; tcgen05.commit multicast frees the stage in both CTAs and publishes the
// accumulator to both); warp 2 = TMEM allocator (both CTAs, cta_group::2); warps 4-7 = top-k epilogue
// over this CTA's 128 accumulator lanes, releasing the accumulator on the leader's barrier.
#include "tvc_internal.h"
#include "tvc_ptx.cuh"
#include "tvc_topk.cuh"

namespace tvc {

namespace {

constexpr int kPStages = 6;
constexpr long long kPaceTimeoutCycles = 400000;   // ~0.2 ms: far beyond any drift pacing is meant to absorb
constexpr int kPABytes = kBM * kBK * 2;        // 128 query rows x 64
constexpr int kPBBytes = 128 * kBK * 2;        // this CTA's 128 of the tile's 256 gallery rows x 64
constexpr int kPStageBytes = kPABytes + kPBBytes;
constexpr int kPSmemStage = kPStages * kPStageBytes;
constexpr int kPSmemBar = kPSmemStage + kStageFloats * 4;
constexpr int kPSmemTotal = kPSmemBar + 256 + 1024;

struct PairBarriers {
  uint64_t full[kPStages];      // used in the leader CTA only
  uint64_t empty[kPStages];     // per CTA, arrived by the leader's multicast commit
  uint64_t tmem_full[2];        // per CTA, arrived by the leader's multicast commit
  uint64_t tmem_empty[2];       // leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t tmem_base;
};

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_topk_pair_kernel(const __grid_constant__ CUtensorMap tmap_q,
                      const __grid_constant__ CUtensorMap tmap_g, const SearchPlan p,
                      float* __restrict__ cand_val, int32_t* __restrict__ cand_idx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* sStage = reinterpret_cast<float*>(smem + kPSmemStage);
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem + kPSmemBar);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_units = plan_units(p);  // m_tiles counts 256-row tiles here

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barrier inits of both CTAs visible before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (threadIdx.x == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    bool pacing = p.pace != nullptr && rank == 0;   // the peer CTA follows through the shared ring
    long long issued = 0;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int split = un.split, mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int q_row = mt * 256 + static_cast<int>(rank) * 128;
      const bool paced = pacing && u < p.full_tiles;
      unsigned int* pace_row = paced ? p.pace + static_cast<size_t>(u / num_pairs) * p.pace_blocks : nullptr;
      for (int nt = t0; nt < t1; ++nt) {
        if (paced && pacing && nt % p.pace_every == 0) {
          const int c = nt / p.pace_every;
          atomicAdd(pace_row + c, 1u);
          if (c >= p.pace_ahead) {
            const volatile unsigned int* behind = pace_row + (c - p.pace_ahead);
            const long long t_start = clock64();
            while (*behind < static_cast<unsigned int>(num_pairs)) {
              if (clock64() - t_start > kPaceTimeoutCycles) {
                pacing = false;   // a pair of this wave is not running: stop waiting for it
                break;
              }
              __nanosleep(256);
            }
          }
        }
        const int g_row = nt * kBN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&bars->full[stage]), 0);
          // profiling only (debug bit 2, results are garbage): the query tile is loaded for the first trip round the
          // ring and stale shared memory multiplied afterwards - what a query operand that never travels would save
          const bool skip_a = (p.debug & 4) && issued >= kPStages;
          ++issued;
          if (rank == 0) mbar_arrive_expect_tx(&bars->full[stage], skip_a ? 2 * kPBBytes : 2 * kPStageBytes);
          uint8_t* sa = smem + stage * kPStageBytes;
          if (!skip_a) tma_load_2d_pair(&tmap_q, full_leader, sa, kb * kBK, q_row, kEvictLast);
          tma_load_2d_pair(&tmap_g, full_leader, sa + kPABytes, kb * kBK, g_row, kEvictNormal);
          if (++stage == kPStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (threadIdx.x == 32 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(256, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int t0 = un.t0, t1 = un.t1;
      for (int nt = t0; nt < t1; ++nt) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          uint8_t* sa = smem + stage * kPStageBytes;
          const uint64_t da = umma_desc_sw128_kmajor(smem_u32(sa));
          const uint64_t db = umma_desc_sw128_kmajor(smem_u32(sa + kPABytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_ss_pair(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                              idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&bars->empty[stage], 3);
          if (++stage == kPStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(&bars->tmem_full[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ top-k epilogue (both CTAs)
    const int q4 = warp & 3;
    const int row_in_tile = q4 * 32 + lane;
    float* my_stage = sStage + row_in_tile;
    TopList<KP> top;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int split = un.split, mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int row = mt * 256 + static_cast<int>(rank) * 128 + row_in_tile;
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
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (row < p.m_rows) topk_store<KP>(top, cand_val, cand_idx,
                                       plan_cand_base(static_cast<long long>(p.full_tiles) * 256, p.splits, KP, row, split));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be signalling our barriers / reading our smem until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int KP>
cudaError_t launch_pair_kp(const CUtensorMap& tq, const CUtensorMap& tg, const SearchPlan& plan, float* cv,
                           int32_t* ci, cudaStream_t stream) {
  static SmemAttrOnce configured;   // per (kernel, device): a second context on another GPU sets its own
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(gemm_topk_pair_kernel<KP>), kPSmemTotal); e != cudaSuccess)
    return e;
  gemm_topk_pair_kernel<KP><<<plan.grid, kThreads, kPSmemTotal, stream>>>(tq, tg, plan, cv, ci);
  note_launch();
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm_topk_pair(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g128,
                                  const SearchPlan& plan, float* cand_val, int32_t* cand_idx,
                                  cudaStream_t stream) {
  switch (plan.kp) {
    case 16:
      return launch_pair_kp<16>(tmap_q, tmap_g128, plan, cand_val, cand_idx, stream);
    case 32:
      return launch_pair_kp<32>(tmap_q, tmap_g128, plan, cand_val, cand_idx, stream);
    case 64:
      return launch_pair_kp<64>(tmap_q, tmap_g128, plan, cand_val, cand_idx, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace tvc
