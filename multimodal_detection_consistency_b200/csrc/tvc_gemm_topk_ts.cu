// Kernel (a), resident-query revision: the fused GEMM + top-k of tvc_gemm_topk_pair.cu with the QUERY operand
// held in tensor memory for the whole work unit instead of being re-staged through shared memory for every
// gallery tile.
//
// Why: a unit multiplies ONE 256-row query tile with thousands of gallery tiles.  The pair kernel streams both
// operands through its shared-memory ring, so half of the L2 -> SM traffic, half of the shared-memory writes and
// half of the shared-memory reads of the tensor cores re-deliver a tile that never changes.  Measured on the
// bench workload with the query loads switched off after the first trip round the ring (debug bit 4, results
// garbage, scripts/perf_probe2.py): SM clock under the 1 kW cap 1410 -> 1485 MHz, kernel 88.4 -> 83.5 ms.  The
// tile does not fit next to a useful ring in shared memory (128 rows x 768 bf16 = 192 KB per CTA), but it fits
// tensor memory: 128 lanes x 384 columns (two bf16 per 32-bit column) of the 512, which leaves 128 columns =
// two 64-column fp32 accumulators.  tcgen05.mma reads A from TMEM ([a_tmem] operand) and B from shared memory.
//
// Shape: cluster of two CTAs, tcgen05 cta_group::2, M = 256 (128 query rows per CTA), N = 64 (32 gallery rows per
// CTA), K = 16.  The planner's gallery tile stays 256 rows = four N = 64 MMA tiles, so plans, pacing counters and
// candidate lists are those of the pair kernel.
// Roles per CTA (256 threads):
//   warp 0 (1 thread)  TMA producer: one 3-D box per stage = this CTA's 32 gallery rows x 4 k-blocks
//                      ([kb][32 rows][64] bf16, 128B-swizzled, 16 KB) into a 12-stage ring (192 KB), completion
//                      bytes on the LEADER's full barrier; paced like the pair kernel.
//   warp 1 (1 thread)  MMA issuer (leader only): waits for the unit's query tile (a_full), then per N = 64 tile
//                      and k-block four tcgen05.mma (A from TMEM columns 128 + 32 kb + 8 k); commit multicast
//                      frees the stage in both CTAs / publishes the accumulator to both.
//   warp 2             TMEM allocator (cta_group::2, 512 columns).
//   warps 4-7          at the start of a unit thread t loads query row t (bf16, d_pad <= 768) from global memory
//                      and stores it into TMEM lane t (tcgen05.st 32x32b.x32, 64 bf16 per instruction), then
//                      arrives on the leader's a_full; afterwards the top-k epilogue of the pair kernel over
//                      64-column accumulators.  The previous unit's MMAs are complete when a warp has consumed that
//                      unit's last accumulator, so the rows can be overwritten; the next unit's MMAs wait on a_full.
// Results are bit-identical to the pair kernel: same bf16 operands, same k order into the same fp32 accumulator,
// same ascending column order into the same top-k list (tests/test_gpu_search.py::test_resident_query_kernel_...).
//
// MEASURED AND NOT USED BY DEFAULT (round 2, profiles/r2j_probe.log): 167 ms against the pair kernel's 88 ms on the
// bench GEMM (81 920 x 1 000 000 x 768), tensor pipe 32 % busy at the full 1965 MHz and 840 W.  An MMA whose A comes
// from tensor memory reads 128 lanes x 32 bytes of it per CTA whatever N is, and that read takes ~50-64 cycles; at
// N = 256 it hides behind the 64 cycles of math, at N = 64 (16 cycles of math) it is the pace.  N = 256 needs 512
// accumulator columns double-buffered, i.e. all of tensor memory, so a resident 768-wide query tile and a full-rate
// MMA exclude each other on this part.  `ts_min_tiles` (tvc_ctx_set_option / TVC_TS_MIN_TILES) switches the kernel on.
#include "tvc_internal.h"
#include "tvc_ptx.cuh"
#include "tvc_topk.cuh"

namespace tvc {

namespace {

constexpr int kTN = 64;                                   // UMMA N: gallery rows per MMA tile (32 per CTA)
constexpr int kTSub = kBN / kTN;                          // MMA tiles per planner tile of kBN gallery rows
constexpr int kTKb = 4;                                   // k-blocks per ring stage (one TMA box)
constexpr int kTSlab = (kTN / 2) * kBK * 2;               // one k-block of this CTA's 32 rows: 4 KB
constexpr int kTStageBytes = kTKb * kTSlab;               // 16 KB
constexpr int kTStages = 12;
constexpr int kTSmemStage = kTStages * kTStageBytes;      // 192 KB
constexpr int kTSmemBar = kTSmemStage + kStageFloats * 4;
constexpr int kTSmemTotal = kTSmemBar + 512 + 1024;
constexpr int kTAccCols = 2 * kTN;                        // accumulators: TMEM columns [0, 128)
constexpr int kTACol0 = kTAccCols;                        // query tile: TMEM columns [128, 128 + 32 * kblocks)
constexpr long long kTPaceTimeoutCycles = 400000;

struct TsBarriers {
  uint64_t full[kTStages];      // used in the leader CTA only
  uint64_t empty[kTStages];     // per CTA, arrived by the leader's multicast commit
  uint64_t tmem_full[2];        // per CTA, arrived by the leader's multicast commit
  uint64_t tmem_empty[2];       // leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint64_t a_full;              // leader only: 8 arrivals per unit (query tile stored in both CTAs' TMEM)
  uint32_t tmem_base;
};

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_topk_ts_kernel(const __grid_constant__ CUtensorMap tmap_g3, const __nv_bfloat16* __restrict__ q_bf,
                    const SearchPlan p, float* __restrict__ cand_val, int32_t* __restrict__ cand_idx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* sStage = reinterpret_cast<float*>(smem + kTSmemStage);
  TsBarriers* bars = reinterpret_cast<TsBarriers*>(smem + kTSmemBar);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_units = plan_units(p);
  const int kstages = (p.kblocks + kTKb - 1) / kTKb;   // ring stages per MMA tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_g3);
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 8);
    }
    mbar_init(&bars->a_full, 8);
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
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int t0 = un.t0, t1 = un.t1;
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
              if (clock64() - t_start > kTPaceTimeoutCycles) {
                pacing = false;   // a pair of this wave is not running: stop waiting for it
                break;
              }
              __nanosleep(256);
            }
          }
        }
#pragma unroll 1
        for (int sub = 0; sub < kTSub; ++sub) {
          const int g_row = nt * kBN + sub * kTN + static_cast<int>(rank) * (kTN / 2);
          for (int ks = 0; ks < kstages; ++ks) {
            mbar_wait(&bars->empty[stage], phase ^ 1u);
            const uint32_t full_leader = mapa_u32(smem_u32(&bars->full[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2 * kTStageBytes);
            // box {64 k, 32 rows, 4 k-blocks}; k-blocks past d_pad / 64 are zero-filled and never multiplied
            tma_load_3d_pair(&tmap_g3, full_leader, smem + stage * kTStageBytes, 0, g_row, ks * kTKb, kEvictNormal);
            if (++stage == kTStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (threadIdx.x == 32 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(256, kTN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0, a_phase = 0;
    const uint32_t a_tmem = tmem_base + static_cast<uint32_t>(kTACol0);
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int t0 = un.t0, t1 = un.t1;
      mbar_wait(&bars->a_full, a_phase);   // this unit's query rows are in both CTAs' tensor memory
      a_phase ^= 1u;
      tc_fence_after();
      for (int nt = t0; nt < t1; ++nt) {
#pragma unroll 1
        for (int sub = 0; sub < kTSub; ++sub) {
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kTN);
          for (int ks = 0; ks < kstages; ++ks) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(smem + stage * kTStageBytes);
#pragma unroll
            for (int j = 0; j < kTKb; ++j) {
              const int kb = ks * kTKb + j;
              if (kb < p.kblocks) {
                const uint64_t db = umma_desc_sw128_kmajor(sb + static_cast<uint32_t>(j * kTSlab));
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)
                  umma_bf16_ts_pair(d_tmem, a_tmem + static_cast<uint32_t>(kb * 32 + k * 8),
                                    db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit_pair(&bars->empty[stage], 3);
            if (++stage == kTStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_pair(&bars->tmem_full[acc], 3);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ query rows -> TMEM, top-k epilogue (both CTAs)
    const int q4 = warp & 3;
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(q4 * 32) << 16;
    float* my_stage = sStage + row_in_tile;
    const size_t row_words = static_cast<size_t>(p.kblocks) * 32;   // 32-bit words per query row (d_pad / 2)
    TopList<KP> top;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int split = un.split, mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int row = mt * 256 + static_cast<int>(rank) * 128 + row_in_tile;
      // (every MMA of the previous unit has completed: this warp consumed that unit's last accumulator)
      {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(q_bf) +
                                                          static_cast<size_t>(row) * row_words);
        for (int c = 0; c < p.kblocks; ++c) {
          uint32_t r[32];
          if (row < p.m_rows) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = __ldg(src + c * 8 + j);
              r[4 * j] = v.x;
              r[4 * j + 1] = v.y;
              r[4 * j + 2] = v.z;
              r[4 * j + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          tmem_st_32x32(tmem_base + lane_bits + static_cast<uint32_t>(kTACol0 + c * 32), r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->a_full), 0));
      }
      const long long self_col = p.skip_self ? static_cast<long long>(row) + p.self_offset : -1ll;
      top.reset();
      float thr = -INFINITY;
      for (int nt = t0; nt < t1; ++nt) {
#pragma unroll 1
        for (int sub = 0; sub < kTSub; ++sub) {
          mbar_wait(&bars->tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + lane_bits + static_cast<uint32_t>(acc * kTN);
          if (!(p.debug & 1))
            topk_consume_tile<KP, kTN>(top, thr, t_addr, my_stage, nt * kBN + sub * kTN, p.n_rows, self_col, p.debug);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
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
cudaError_t launch_ts_kp(const CUtensorMap& tg3, const __nv_bfloat16* q_bf, const SearchPlan& plan, float* cv,
                         int32_t* ci, cudaStream_t stream) {
  static SmemAttrOnce configured;   // per (kernel, device)
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(gemm_topk_ts_kernel<KP>), kTSmemTotal); e != cudaSuccess)
    return e;
  gemm_topk_ts_kernel<KP><<<plan.grid, kThreads, kTSmemTotal, stream>>>(tg3, q_bf, plan, cv, ci);
  note_launch();
  return cudaGetLastError();
}

}  // namespace

int ts_max_kblocks() { return (512 - kTAccCols) / 32; }

cudaError_t launch_gemm_topk_ts(const CUtensorMap& tmap_g3, const __nv_bfloat16* q_bf, const SearchPlan& plan,
                                float* cand_val, int32_t* cand_idx, cudaStream_t stream) {
  if (plan.kblocks > ts_max_kblocks() || !plan.pair) return cudaErrorInvalidValue;
  switch (plan.kp) {
    case 16:
      return launch_ts_kp<16>(tmap_g3, q_bf, plan, cand_val, cand_idx, stream);
    case 32:
      return launch_ts_kp<32>(tmap_g3, q_bf, plan, cand_val, cand_idx, stream);
    case 64:
      return launch_ts_kp<64>(tmap_g3, q_bf, plan, cand_val, cand_idx, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace tvc
