// Internal declarations shared by the libtvc.so translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/tvc.h"

namespace tvc {

// ---- tile geometry of the fused GEMM + top-k kernel (tvc_gemm_topk.cu) ----------------------
constexpr int kBM = 128;      // query rows per CTA tile (UMMA M, one TMEM lane per row)
constexpr int kBN = 256;      // gallery rows per MMA tile (UMMA N)
constexpr int kBK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kStages = 4;    // smem ring depth
constexpr int kABytes = kBM * kBK * 2;
constexpr int kBBytes = kBN * kBK * 2;
constexpr int kThreads = 256;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue

struct SearchPlan {
  int m_rows;           // valid query rows
  int n_rows;           // valid gallery rows
  int kblocks;          // d_pad / 64
  int m_tiles;
  int n_tiles;
  // Two kinds of work unit.  The first `full_tiles` query tiles (whole waves of the persistent grid) each run
  // against the WHOLE gallery (one candidate list per row); the remaining `rem_tiles` query tiles - less than
  // one wave - are cut into `splits` gallery ranges of `tiles_per_split` tiles so that they fill the last wave
  // (`splits` lists per row, merged by the re-rank).  Long units matter: a row's running threshold only
  // becomes selective after ~16k gallery columns, and until then nearly every 32 x 32 chunk takes the
  // epilogue's slow path (75 % of the chunks of a 163-tile unit, 8 % of a 3900-tile one).
  int full_tiles;
  int rem_tiles;
  int splits;           // gallery ranges per query tile of the remainder (each owns a candidate list)
  int tiles_per_split;
  int kp;               // candidates kept per (row, split): 16 / 32 / 64
  int grid;
  int pair;             // 1: CTA-pair kernel (256-row query tiles, cta_group::2)
  int debug;            // profiling only: bit0 = skip the top-k epilogue (results are garbage)
  int skip_self;        // drop candidate == query row (+ self_offset)
  int64_t self_offset;  // global row of query 0 minus global row of gallery row 0
  // Pacing of the CTA pairs of a wave (pair kernel, whole-gallery units only; null = off).  The pairs of a wave
  // stream the same gallery tiles; left alone they drift apart by more than the L2 holds and the stragglers
  // fetch every tile from HBM again (ncu, round 1: 22.8 GB per launch against a 7.1 GB floor).  pace[w *
  // pace_blocks + c] counts the pairs of wave w that have reached gallery tile c * pace_every; a producer enters
  // block c only once every pair of its wave has reached block c - pace_ahead, so the spread between the first
  // and the last pair of a wave stays below pace_ahead * pace_every tiles.  A hint, not a barrier: a wait that
  // times out (a pair that is not resident because another kernel holds its SMs) switches pacing off for that CTA.
  unsigned int* pace;
  int pace_every;
  int pace_ahead;
  int pace_blocks;
};

// plans the unit decomposition for `sm_count` persistent CTAs
SearchPlan make_search_plan(int64_t m, int64_t n, int d_pad, int k, int sm_count, bool pair);

// unit u -> (query tile, gallery range, tile interval); candidate list of (row, range)
struct SearchUnit {
  int mt, split, t0, t1;
};
__host__ __device__ inline int plan_units(const SearchPlan& p) { return p.full_tiles + p.rem_tiles * p.splits; }
__host__ __device__ inline SearchUnit plan_unit(const SearchPlan& p, int u) {
  SearchUnit x;
  if (u < p.full_tiles) {
    x.mt = u;
    x.split = 0;
    x.t0 = 0;
    x.t1 = p.n_tiles;
  } else {
    const int j = u - p.full_tiles;
    x.split = j / p.rem_tiles;
    x.mt = p.full_tiles + (j - x.split * p.rem_tiles);
    x.t0 = x.split * p.tiles_per_split;
    x.t1 = x.t0 + p.tiles_per_split < p.n_tiles ? x.t0 + p.tiles_per_split : p.n_tiles;
  }
  return x;
}
// rows below full_rows = full_tiles * (rows per query tile) own one list, the others `splits` lists
__host__ __device__ inline size_t plan_cand_base(long long full_rows, int splits, int kp, long long row, int split) {
  return row < full_rows ? static_cast<size_t>(row) * kp
                         : static_cast<size_t>(full_rows) * kp +
                               (static_cast<size_t>(row - full_rows) * splits + split) * kp;
}

cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g,
                             const SearchPlan& plan, float* cand_val, int32_t* cand_idx,
                             cudaStream_t stream);

// CTA-pair revision (tvc_gemm_topk_pair.cu); tmap_g128 has a 128-row box (each CTA stages half a tile)
cudaError_t launch_gemm_topk_pair(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g128,
                                  const SearchPlan& plan, float* cand_val, int32_t* cand_idx,
                                  cudaStream_t stream);

// pair kernel with the first 7 k-blocks of the query tile resident in shared memory for the whole unit
cudaError_t launch_gemm_topk_pair_rq(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g128,
                                     const SearchPlan& plan, float* cand_val, int32_t* cand_idx, int resident,
                                     cudaStream_t stream);

// resident-query revision (tvc_gemm_topk_ts.cu): the query tile lives in tensor memory for the whole unit;
// tmap_g3 = the gallery as {64, rows, k-blocks} with a {64, 32, 4} box; q_bf = prepared bf16 queries [m, d_pad]
int ts_max_kblocks();
cudaError_t launch_gemm_topk_ts(const CUtensorMap& tmap_g3, const __nv_bfloat16* q_bf, const SearchPlan& plan,
                                float* cand_val, int32_t* cand_idx, cudaStream_t stream);

cudaError_t launch_gemm_store(const CUtensorMap& tmap_q, const CUtensorMap& tmap_g, int m, int n,
                              int kblocks, float* out, int64_t ld_out, int sm_count,
                              cudaStream_t stream);

// ---- auxiliary kernels (tvc_aux.cu) -----------------------------------------------------------
// rows [n, d] of dtype -> bf16 [n, d_pad] (zero padded) and optional fp32 [n, d]; optional L2 norm.
cudaError_t launch_prep_rows(const void* rows, int dtype, int64_t n, int d, int d_pad, bool normalize,
                             __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t stream);

// same, with the bf16 rows written at row `dst_row0` of up to kMaxParts destinations (own / peer query buffers)
constexpr int kMaxParts = 16;
struct BcastSpec {
  int n;
  __nv_bfloat16* bf16[kMaxParts];
};
cudaError_t launch_prep_rows_bcast(const void* rows, int dtype, int64_t n, int d, int d_pad, bool normalize,
                                   const BcastSpec& dst, int64_t dst_row0, float* out_f32, cudaStream_t stream);

// select top-kp by GEMM score over splits, re-score in fp32 from the masters, emit ordered top-k.
cudaError_t launch_rerank(const float* cand_val, const int32_t* cand_idx, int64_t m, int64_t full_rows, int splits,
                          int kp, int k, const float* q_f32, const float* g_f32, int d,
                          float threshold, int64_t global_row_offset, float* out_sim,
                          int64_t* out_idx, cudaStream_t stream);

cudaError_t launch_merge_topk(const float* in_sim, const int64_t* in_idx, int64_t m, int parts, int k,
                              float* out_sim, int64_t* out_idx, cudaStream_t stream);

// part_scratch: nullptr = single-pass kernels; otherwise k_occurrence_part_scratch_bytes(...) bytes (256-byte aligned)
// for the bucketed two-pass path.  That function returns 0 when the path does not apply (stream shorter than
// part_min entries, histogram that fits shared memory or exceeds 4 Mi bins, misaligned stream).
size_t k_occurrence_part_scratch_bytes(const int64_t* idx, int64_t m, int k, int64_t n_bins, int64_t part_min);
cudaError_t launch_k_occurrence(const int64_t* idx, int64_t m, int k, int64_t idx_base,
                                int64_t n_bins, int32_t* counts, int sm_count, int* flag_scratch,
                                void* part_scratch, cudaStream_t stream);

struct MetricKs {
  int n;
  int k[8];
};
cudaError_t launch_retrieval_metrics(const int64_t* topk, int64_t nq, int k, const int64_t* rel_ptr,
                                     const int64_t* rel_idx, const MetricKs& ks, float* out, cudaStream_t stream);

cudaError_t launch_gather_rows(const float* g_f32, const __nv_bfloat16* g_bf16, int d, int d_pad,
                               const int64_t* idx, int64_t n, int64_t n_rows, float* out,
                               cudaStream_t stream);

// Where the rows behind global indices live: up to kMaxParts row shards, each resident in this GPU's
// HBM or in a PEER GPU's HBM mapped through CUDA IPC (read over NVLink by the kernel itself).
struct RowSource {
  int nparts;
  int d_pad;
  const float* f32[kMaxParts];            // fp32 master of shard p (or null)
  const __nv_bfloat16* bf16[kMaxParts];   // bf16 rows of shard p (used when there is no master)
  long long off[kMaxParts];               // global index of the shard's first row
  long long n[kMaxParts];                 // rows in the shard
};

// Destination of the candidates of a sharded search: row r belongs to query slice r / rows_per_slice,
// whose owner's receive buffers [n_slices, rows_in_slice, kp] (local or peer-mapped) are val/idx[j].
struct ScatterSpec {
  int n_slices;            // 0: no scatter, write the local [m, kp] lists
  int slot;                // this shard's position in every receive buffer
  long long rows_per_slice;
  float* val[kMaxParts];
  long long* idx[kMaxParts];
};

cudaError_t launch_select_candidates(const float* cand_val, const int32_t* cand_idx, int64_t m, int64_t full_rows,
                                     int splits, int kp, int64_t global_row_offset, const ScatterSpec& sc, float* out_val,
                                     int64_t* out_idx, cudaStream_t stream);

// re-score at the shards (exchange_* kernels, tvc_aux.cu)
struct ReqDst {
  int n;
  long long* req[kMaxParts];     // per shard: this owner's block [rows_in_slice, kp] inside the shard's request area
};
struct ScoreDst {
  float* score[kMaxParts];       // per owner: its score area [rows_in_slice, kp]
};
cudaError_t launch_exchange_merge(const float* cand_val, const int64_t* cand_idx, int64_t m, int parts, int kp,
                                  const ReqDst& dst, cudaStream_t stream);
cudaError_t launch_exchange_rescore(const int64_t* req, const float* q_f32, const float* g_f32, int64_t g_off,
                                    int64_t g_n, int d, int owners, int64_t rows_per_slice, int64_t m_total, int kp,
                                    const ScoreDst& dst, cudaStream_t stream);
cudaError_t launch_exchange_finalize(const int64_t* req, const float* score, int64_t m, int kp, int k, float threshold,
                                     float* out_sim, int64_t* out_idx, cudaStream_t stream);

cudaError_t launch_rerank_merged(const float* cand_val, const int64_t* cand_idx, int64_t m, int parts, int kp,
                                 int k, const float* q_f32, const RowSource& src, int d, float threshold,
                                 float* out_sim, int64_t* out_idx, cudaStream_t stream);

struct ConsistencyEmbArgs {
  const float* img;
  const float* txt;
  const float* var;
  RowSource ret;          // retrieval gallery shards
  const int64_t* ret_idx;
  int n_ret_cand;
  const float* gen;  // direct generative embeddings [Q,G,d] (or null)
  const int32_t* g_cnt;
  RowSource genr;         // generative (bank) shards
  const int64_t* gen_idx;
  int n_gen_cand;
  float* out_sv;
  float* out_sr;
  float* out_sg;
  // similarity lists written by the pipelined kernel for the statistics kernel (device workspace)
  float* w_s0;
  float* w_sv;
  float* w_sr;
  int32_t* w_rcnt;
  float* w_sg;
  int32_t* w_gcnt;
  float* w_sx;
  unsigned long long* trace;   // debugging: per-query timestamps of block 0 (null = off)
};

cudaError_t launch_consistency_sims(const tvc_detector_params& p, int64_t q, const float* s0,
                                    const float* sv, const float* sr, const int32_t* r_cnt,
                                    const float* sg, const int32_t* g_cnt, const float* sxv,
                                    float* scores, uint8_t* flags, cudaStream_t stream);

cudaError_t launch_consistency_emb(const tvc_detector_params& p, int64_t q, int d,
                                   const ConsistencyEmbArgs& a, float* scores, uint8_t* flags,
                                   int sm_count, int force_generic, cudaStream_t stream);

cudaError_t launch_reference_vector(int64_t q, int d, int v, const float* img, const RowSource& src,
                                    const int64_t* ret_idx, int k, const float* gen, int m, float sigma_threshold,
                                    float* out_s, float* out_ref, float* out_sigma, uint8_t* flags,
                                    uint8_t* valid_ws, int sm_count, cudaStream_t stream);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, DEVICE): one flag per process would
// leave the kernel unconfigured on the second GPU a process opens a context on.  One instance per call site.
struct SmemAttrOnce {
  std::atomic<uint64_t> done{0};   // bit = device ordinal (mod 64)
  cudaError_t ensure(const void* kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// counts every kernel launch made by the library (reported through tvc_ctx_launch_count)
void note_launch(int n = 1);
int64_t launches_so_far();

}  // namespace tvc
