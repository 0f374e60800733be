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

  // Warp index and CTA rank in forms the compiler can prove warp-uniform (a shuffle result, blockIdx): the two
  // issue loops below then run on the uniform datapath.  As single-thread branches (threadIdx.x == 0 / 32) every
  // UTCHMMA / UTMALDG sat in a divergence "waterfall" (ELECT + 5 R2UR + BRA.U.ANY each) and one k-block cost the
  // issuing thread ~90 dependent instructions - about the 512 cycles the tensor cores need for it, so the issue
  // thread, not the tensor pipe, set the pace (a handful of extra instructions in that loop cost 6 %).
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;         // == %cluster_ctarank of a (2,1,1) cluster; 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_units = plan_units(p);  // m_tiles counts 256-row tiles here
  // profiling only: debug bits 4-6 = use fewer stages of the ring (how thin a ring still covers the L2 latency)
  const int nstages = ((p.debug >> 4) & 7) ? min(kPStages, (p.debug >> 4) & 7) : kPStages;

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
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars_base = smem_u32(bars);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs; whole warp
    // walks the loop, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    bool pacing = p.pace != nullptr && rank == 0;   // the peer CTA follows through the shared ring
    long long issued = 0;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int q_row = mt * 256 + static_cast<int>(rank) * 128;
      const bool paced = pacing && u < p.full_tiles;
      unsigned int* pace_row = paced ? p.pace + static_cast<size_t>(u / num_pairs) * p.pace_blocks : nullptr;
      for (int nt = t0; nt < t1; ++nt) {
        if (paced && pacing && nt % p.pace_every == 0) {
          const int c = nt / p.pace_every;
          int keep = 1;
          if (lane == 0) {
            atomicAdd(pace_row + c, 1u);
            if (c >= p.pace_ahead) {
              const volatile unsigned int* behind = pace_row + (c - p.pace_ahead);
              const long long t_start = clock64();
              while (*behind < static_cast<unsigned int>(num_pairs)) {
                if (clock64() - t_start > kPaceTimeoutCycles) {
                  keep = 0;   // a pair of this wave is not running: stop waiting for it
                  break;
                }
                __nanosleep(256);
              }
            }
          }
          pacing = __shfl_sync(0xffffffffu, keep, 0) != 0;
        }
        const int g_row = nt * kBN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          // the leader's barrier of this stage: same offset, CTA-rank bit of the shared::cluster address cleared
          const uint32_t full_leader = (bars_base + static_cast<uint32_t>(offsetof(PairBarriers, full) + 8 * stage)) & kLeaderCtaMask;
          // profiling only (debug bit 2, results are garbage): the query tile is loaded for the first trip round the
          // ring and stale shared memory multiplied afterwards - what a query operand that never travels would save
          const bool skip_a = (p.debug & 4) && issued >= kPStages;
          ++issued;
          const uint32_t sa = smem_base + static_cast<uint32_t>(stage * kPStageBytes);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&bars->full[stage], skip_a ? 2 * kPBBytes : 2 * kPStageBytes);
            if (!skip_a) tma_load_2d_pair_u32(&tmap_q, full_leader, sa, kb * kBK, q_row, kEvictLast);
            tma_load_2d_pair_u32(&tmap_g, full_leader, sa + kPABytes, kb * kBK, g_row, kEvictNormal);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader only; whole warp walks
    // the loop, one elected lane - always the same - issues the MMAs and their commits)
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
          const uint32_t sa = smem_base + static_cast<uint32_t>(stage * kPStageBytes);
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sa + kPABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_ss_pair(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&bars->empty[stage], 3);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit_pair(&bars->tmem_full[acc], 3);
        __syncwarp();
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

// ------------------------------------------------------------------------------------------------------------
// Pair kernel with a (mostly) RESIDENT query tile.  A unit multiplies one 256-row query tile with thousands of
// gallery tiles; the kernel above re-delivers that tile through the ring for every gallery tile - half of the
// L2 -> SM traffic and of the shared-memory writes.  Not re-loading it at all is worth +5.9 % under the 1 kW cap
// (stale-operand probe, debug bit 4, profiles/r2i_probe_a.log); the ring needs only 3 k-blocks of lookahead to
// cover the L2 latency (4 of the 6 stages: same speed, 3: -11 %, profiles/r2k_ring.log).  So here the first
// here 8 of the k-blocks of this CTA's 128 query rows (128 KB), spread evenly over the row, stay in shared
// memory for the whole unit, and the ring becomes 6 slots of 16 KB that carry the gallery k-blocks and the query
// k-blocks that did not fit (the epilogue's slow path stages in thread-local memory to make room).  Same MMAs on the same operands in the same order as the kernel
// above: bit-identical results.
//   warp 0    TMA producer of the ring (gallery k-blocks, streamed query k-blocks), paced
//   warp 1    MMA issuer (leader): per unit waits a_full; A descriptors point into the resident area or a slot;
//             after a unit's last MMA a multicast commit on a_empty lets both CTAs replace the resident tile
//   warp 3    resident-tile loader: waits a_empty (previous unit computed), loads the next unit's k-blocks,
//             completion on the leader's a_full - on its own warp so that the ring keeps streaming meanwhile
// (all three as warp-uniform loops with one elected lane issuing)
constexpr int kRqUnits = 14;                                 // 16 KB units of shared memory: resident + ring slots
                                                             // (the epilogue's slow path stages in local memory)
constexpr int kRqMaxSlots = 8;
constexpr int kRqSlotBytes = kPABytes;                       // 16 KB: [128 rows x 64] bf16, A or B
constexpr int kRqSmemRing = kRqUnits * kRqSlotBytes;         // end of the operand area (208 KB)
constexpr int kRqSmemBar = kRqSmemRing;
constexpr int kRqSmemTotal = kRqSmemBar + 256 + 1024;
static_assert(kPABytes == kPBBytes, "one slot size for both operands");

struct RqBarriers {
  uint64_t full[kRqMaxSlots];   // used in the leader CTA only
  uint64_t empty[kRqMaxSlots];  // per CTA, arrived by the leader's multicast commit
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a_full;              // leader only: the unit's resident query k-blocks of both CTAs have landed
  uint64_t a_empty;             // per CTA: every MMA of the unit has completed
  uint32_t tmem_base;
};

template <int KP, int kRqResident>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_topk_pair_rq_kernel(const __grid_constant__ CUtensorMap tmap_q,
                         const __grid_constant__ CUtensorMap tmap_g, const SearchPlan p,
                         float* __restrict__ cand_val, int32_t* __restrict__ cand_idx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kRqSlots = kRqUnits - kRqResident;
  constexpr int kRqSmemRes = kRqResident * kRqSlotBytes;
  static_assert(kRqSlots <= kRqMaxSlots && kRqSlots >= 4, "ring size");   // resident 6 / 7 / 8 -> 8 / 7 / 6 slots
  RqBarriers* bars = reinterpret_cast<RqBarriers*>(smem + kRqSmemBar);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;         // == %cluster_ctarank of a (2,1,1) cluster; 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_units = plan_units(p);
  const int n_res = p.kblocks < kRqResident ? p.kblocks : kRqResident;   // resident query k-blocks
  // WHICH k-blocks stay: spread evenly over the row (bit kb of res_mask), so that the streamed ones - two ring slots
  // each instead of one - never come back to back and the ring's lookahead stays even (contiguous: 6 resident of 12
  // was slower than 5, profiles/r2v_probe.log).  Resident k-block kb lives in unit popc(res_mask below bit kb).
  uint32_t res_mask = 0;
  for (int kb = 0; kb < p.kblocks; ++kb)
    if ((kb + 1) * n_res / p.kblocks > kb * n_res / p.kblocks) res_mask |= 1u << kb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kRqSlots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 8);
    }
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t res_base = smem_u32(smem);                    // resident k-block kb at res_base + kb * 16 KB
  const uint32_t ring_base = res_base + kRqSmemRes;            // ring slot s at ring_base + s * 16 KB
  const uint32_t bars_base = smem_u32(bars);

  if (warp == 0) {
    // ------------------------------------------------------------------ ring producer (both CTAs)
    int slot = 0;
    uint32_t phase = 0;
    bool pacing = p.pace != nullptr && rank == 0;
    auto push = [&](const CUtensorMap* tm, int c0, int c1, uint64_t hint) {
      mbar_wait(&bars->empty[slot], phase ^ 1u);
      const uint32_t full_leader = (bars_base + static_cast<uint32_t>(offsetof(RqBarriers, full) + 8 * slot)) & kLeaderCtaMask;
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&bars->full[slot], 2 * kRqSlotBytes);
        tma_load_2d_pair_u32(tm, full_leader, ring_base + static_cast<uint32_t>(slot * kRqSlotBytes), c0, c1, hint);
      }
      __syncwarp();
      if (++slot == kRqSlots) {
        slot = 0;
        phase ^= 1u;
      }
    };
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int mt = un.mt, t0 = un.t0, t1 = un.t1;
      const int q_row = mt * 256 + static_cast<int>(rank) * 128;
      const bool paced = pacing && u < p.full_tiles;
      unsigned int* pace_row = paced ? p.pace + static_cast<size_t>(u / num_pairs) * p.pace_blocks : nullptr;
      for (int nt = t0; nt < t1; ++nt) {
        if (paced && pacing && nt % p.pace_every == 0) {
          const int c = nt / p.pace_every;
          int keep = 1;
          if (lane == 0) {
            atomicAdd(pace_row + c, 1u);
            if (c >= p.pace_ahead) {
              const volatile unsigned int* behind = pace_row + (c - p.pace_ahead);
              const long long t_start = clock64();
              while (*behind < static_cast<unsigned int>(num_pairs)) {
                if (clock64() - t_start > kPaceTimeoutCycles) {
                  keep = 0;
                  break;
                }
                __nanosleep(256);
              }
            }
          }
          pacing = __shfl_sync(0xffffffffu, keep, 0) != 0;
        }
        const int g_row = nt * kBN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (!((res_mask >> kb) & 1u)) push(&tmap_q, kb * kBK, q_row, kEvictLast);
          push(&tmap_g, kb * kBK, g_row, kEvictNormal);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ resident-tile loader (both CTAs)
    uint32_t ph = 0;
    bool first = true;
    const uint32_t a_full_leader = (bars_base + static_cast<uint32_t>(offsetof(RqBarriers, a_full))) & kLeaderCtaMask;
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int q_row = un.mt * 256 + static_cast<int>(rank) * 128;
      if (!first) {
        mbar_wait_parked(&bars->a_empty, ph);   // the previous unit's MMAs have read the tile for the last time (a
        ph ^= 1u;                               // whole unit away: sleep between polls)
      }
      first = false;
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&bars->a_full, static_cast<uint32_t>(2 * n_res * kRqSlotBytes));
        int unit = 0;
        for (int kb = 0; kb < p.kblocks; ++kb)
          if ((res_mask >> kb) & 1u) {
            tma_load_2d_pair_u32(&tmap_q, a_full_leader, res_base + static_cast<uint32_t>(unit * kRqSlotBytes), kb * kBK,
                                 q_row, kEvictNormal);
            ++unit;
          }
      }
      __syncwarp();
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(256, kBN);
    int slot = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0, a_phase = 0;
    auto advance = [&]() {
      if (++slot == kRqSlots) {
        slot = 0;
        phase ^= 1u;
      }
    };
    auto mma4 = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool first_kb, int free_a_slot, int free_b_slot) {
      const uint64_t da = umma_desc_sw128_kmajor(a_addr);
      const uint64_t db = umma_desc_sw128_kmajor(b_addr);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16_ss_pair(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                            (first_kb && k == 0) ? 0u : 1u);
        if (free_a_slot >= 0) umma_commit_pair(&bars->empty[free_a_slot], 3);
        umma_commit_pair(&bars->empty[free_b_slot], 3);
      }
      __syncwarp();
    };
    for (int u = pair; u < total_units; u += num_pairs) {
      const SearchUnit un = plan_unit(p, u);
      const int t0 = un.t0, t1 = un.t1;
      mbar_wait(&bars->a_full, a_phase);
      a_phase ^= 1u;
      tc_fence_after();
      for (int nt = t0; nt < t1; ++nt) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if ((res_mask >> kb) & 1u) {                // query k-block resident, gallery k-block from the ring
            const int unit = __popc(res_mask & ((1u << kb) - 1u));
            mbar_wait(&bars->full[slot], phase);
            tc_fence_after();
            mma4(d_tmem, res_base + static_cast<uint32_t>(unit * kRqSlotBytes),
                 ring_base + static_cast<uint32_t>(slot * kRqSlotBytes), kb == 0, -1, slot);
            advance();
          } else {                                    // both from the ring: query slot, then gallery slot
            mbar_wait(&bars->full[slot], phase);
            const int a_slot = slot;
            advance();
            mbar_wait(&bars->full[slot], phase);
            tc_fence_after();
            mma4(d_tmem, ring_base + static_cast<uint32_t>(a_slot * kRqSlotBytes),
                 ring_base + static_cast<uint32_t>(slot * kRqSlotBytes), kb == 0, a_slot, slot);
            advance();
          }
        }
        if (elect_one()) umma_commit_pair(&bars->tmem_full[acc], 3);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (elect_one()) umma_commit_pair(&bars->a_empty, 3);   // both CTAs may now replace their resident k-blocks
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ top-k epilogue (both CTAs)
    const int q4 = warp & 3;
    const int row_in_tile = q4 * 32 + lane;
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
        if (!(p.debug & 1))
          topk_consume_tile<KP, kBN, true>(top, thr, t_addr, nullptr, nt * kBN, p.n_rows, self_col, p.debug);
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
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int KP, int RES>
cudaError_t launch_pair_rq_kp(const CUtensorMap& tq, const CUtensorMap& tg, const SearchPlan& plan, float* cv,
                              int32_t* ci, cudaStream_t stream) {
  static SmemAttrOnce configured;
  if (cudaError_t e = configured.ensure(reinterpret_cast<const void*>(gemm_topk_pair_rq_kernel<KP, RES>), kRqSmemTotal);
      e != cudaSuccess)
    return e;
  gemm_topk_pair_rq_kernel<KP, RES><<<plan.grid, kThreads, kRqSmemTotal, stream>>>(tq, tg, plan, cv, ci);
  note_launch();
  return cudaGetLastError();
}

template <int KP>
cudaError_t launch_pair_rq_res(const CUtensorMap& tq, const CUtensorMap& tg, const SearchPlan& plan, float* cv,
                               int32_t* ci, int resident, cudaStream_t stream) {
  switch (resident) {
    case 9:
      return launch_pair_rq_kp<KP, 9>(tq, tg, plan, cv, ci, stream);
    case 8:
      return launch_pair_rq_kp<KP, 8>(tq, tg, plan, cv, ci, stream);
    case 7:
      return launch_pair_rq_kp<KP, 7>(tq, tg, plan, cv, ci, stream);
    default:
      return cudaErrorInvalidValue;
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

cudaError_t launch_gemm_topk_pair_rq(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g128,
                                     const SearchPlan& plan, float* cand_val, int32_t* cand_idx, int resident,
                                     cudaStream_t stream) {
  // 14 units of 16 KB (the epilogue's slow path stages in thread-local memory, so no shared-memory staging
  // buffer): `resident` query k-blocks (spread evenly over the row) + a ring of the rest.  Measured on the bench
  // GEMM against the plain pair kernel (profiles/r3c_probe.log, r3d_probe.log): d = 768: 6 / 7 / 8 / 9 resident =
  // +3.1 / +3.3 / +4.0 / +3.0 %; d = 512: +4.5 / +5.5 / +6.7 %.  A streamed k-block takes two ring slots, and the ring
  // must still hold ~2 k-blocks of prefetch on top of the 2 inside the tensor pipe: 6 slots are the least that do.
  if (resident <= 0) resident = 8;   // (ring: 14 - resident slots)
  switch (plan.kp) {
    case 16:
      return launch_pair_rq_res<16>(tmap_q, tmap_g128, plan, cand_val, cand_idx, resident, stream);
    case 32:
      return launch_pair_rq_res<32>(tmap_q, tmap_g128, plan, cand_val, cand_idx, resident, stream);
    case 64:
      return launch_pair_rq_res<64>(tmap_q, tmap_g128, plan, cand_val, cand_idx, resident, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace tvc
