/*
 * tvc.h — C ABI of libtvc.so: the B200 (sm_100a) text-variant-consistency (TVC) scoring and
 * retrieval hot path.
 *
 * Every entry point replaces a library call site of the reference (which is 100 % Python and has
 * no FFI of its own; citations are path:line under the reference tree):
 *
 *   tvc_gallery_create / _append   faiss.IndexFlatIP(d) + index.add(features)   src/retrieval.py:477-525
 *                                  np.array([ref.vector ...]) rebuilt per query  src/ref_bank.py:475
 *   tvc_search                     index.search(q.astype(f32), k)               src/retrieval.py:652-656
 *                                  cosine_similarity + np.argsort[::-1][:k]      src/retrieval.py:669-671
 *                                  dot/(|r||q|+1e-8), where(>=thr), argsort      src/ref_bank.py:475-484,197-203
 *                                  np.dot + np.argpartition                      experiments/defenses/retrieval_ref.py:246-290
 *                                  F.cosine_similarity + torch.topk(.,1)         src/attacks/hubness_attack.py:482-489
 *   tvc_similarity_matrix          cosine_similarity(T, I) / np.dot(T, I.T)      src/retrieval.py:706-708
 *                                  torch.mm(x_hat, y_hat.T)                      src/utils/metrics.py:156-159
 *   tvc_consistency_sims / _emb    per-pair cosine loops + mean/std/var/min/max  src/detector.py:461-485,528-542,573-579,653-682,399
 *                                                                                experiments/defenses/detector.py:228-300
 *                                                                                experiments/defenses/consistency_checker.py:74-272
 *                                                                                experiments/defenses/text_variants.py:412-451
 *   tvc_reference_vector_rule      the documented rule: mean reference vectors, S = cos, sigma = std(S)  README.md:474-482,846
 *   tvc_k_occurrence               hubness_counts[j] += 1 double loop            references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:43-57
 *                                  (top1 == 0).sum()                             src/attacks/hubness_attack.py:492-496
 *   tvc_retrieval_metrics          argsort + per-query Python loops (Recall/Precision/NDCG@K, RR, AP)  src/utils/metrics.py:386-574
 *   tvc_merge_topk                 (new) merge of per-shard top-k candidates after the NCCL all-gather
 *   tvc_search_candidates /        (new) sharded search: candidates written into the owner's HBM over NVLink,
 *   tvc_rerank_candidates                re-ranked there from local + peer fp32 masters
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a tvc_status (0 = ok) and never throws.
 *   - data pointers may be host or device memory (detected with cudaPointerGetAttributes); host
 *     buffers are staged through the context's device workspace and the call returns after the
 *     results are back in the host buffer. With device pointers the call is asynchronous on `stream`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - search results are ordered by (similarity descending, index ascending); unused slots carry
 *     index -1 and similarity -inf (the FAISS convention the reference relies on,
 *     experiments/defenses/retrieval_ref.py:257).
 *   - a gallery is immutable under concurrent tvc_search calls; tvc_gallery_append needs external
 *     exclusion (the Python wrappers hold the lock the reference's ReferenceBank holds).
 */
#ifndef TVC_H_
#define TVC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVC_VERSION 100

typedef enum {
  TVC_OK = 0,
  TVC_ERR_INVALID = 1,     /* bad argument (null pointer, d mismatch, k out of range ...) */
  TVC_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed; see tvc_last_error */
  TVC_ERR_NO_DEVICE = 3,   /* no sm_100 device: there is no CPU fallback */
  TVC_ERR_UNSUPPORTED = 4, /* valid request this build cannot serve (k > TVC_MAX_K ...) */
  TVC_ERR_OOM = 5
} tvc_status;

typedef enum { TVC_F32 = 0, TVC_BF16 = 1, TVC_F16 = 2 } tvc_dtype;

/* gallery flags */
#define TVC_GALLERY_NORMALIZE 1u   /* L2-normalise rows at ingest (cosine metric; ReferenceBank) */
#define TVC_GALLERY_NO_MASTER 2u   /* keep bf16 rows only: no fp32 master, no fp32 re-rank */

/* search flags */
#define TVC_SEARCH_NORMALIZE_Q 1u  /* L2-normalise query rows (cosine metric) */
#define TVC_SEARCH_SKIP_SELF 2u    /* query set == gallery: drop candidate idx == query row (hubness spec [:,1:k+1]) */
#define TVC_SEARCH_PREPARED_Q 4u   /* tvc_search_candidates only: `queries` is the bf16 [m, tvc_query_row_bytes(d)/2]
                                      operand written by tvc_prepare_queries (q_dtype is ignored) */

#define TVC_MAX_K 56               /* largest k served by the in-register top-k epilogue */
#define TVC_MAX_VARIANTS 16
#define TVC_MAX_REFS 16
#define TVC_NSCORES 24             /* floats written per query by tvc_consistency_* */

typedef struct tvc_ctx tvc_ctx;
typedef struct tvc_gallery tvc_gallery;

/* Column layout of the [Q, TVC_NSCORES] score matrix written by tvc_consistency_*. */
enum {
  TVC_S_ORIGINAL = 0,        /* s0 = cos(image, text)                         detector.py:461 / defenses/detector.py:240 */
  TVC_S_TV_MEAN = 1,         /* text_variant_consistency = mean_v s_v          defenses/detector.py:252 */
  TVC_S_TV_STD = 2,          /* text_variant_std (ddof=0)                      defenses/detector.py:253 */
  TVC_S_TV_MIN = 3,
  TVC_S_TV_VAR = 4,
  TVC_S_RET_MEAN = 5,        /* retrieval_consistency                          defenses/detector.py:266 */
  TVC_S_RET_STD = 6,
  TVC_S_GEN_MEAN = 7,        /* generative_consistency / mean_similarity       defenses/detector.py:279, detector.py:537 */
  TVC_S_GEN_STD = 8,
  TVC_S_GEN_MAX = 9,         /* max_similarity                                 detector.py:538 */
  TVC_S_CROSS_MODAL_VAR = 10,/* var(ddof=0) of the positive means              defenses/detector.py:295-300 */
  TVC_S_XV_MEAN = 11,        /* variant<->variant pair cosines                 defenses/text_variants.py:412-451 */
  TVC_S_XV_MIN = 12,
  TVC_S_XV_VAR = 13,
  TVC_S_DET_TV = 14,         /* 1-(0.7*consistency+0.3*variability)            detector.py:474-485 */
  TVC_S_DET_SD = 15,         /* 1-mean(r)                                      detector.py:542 */
  TVC_S_DET_C = 16,          /* 1-s0                                           detector.py:573-579 */
  TVC_S_DET_AGG = 17,        /* _aggregate_scores                              detector.py:643-682 */
  TVC_S_CC_OVERALL = 18,     /* ConsistencyChecker overall score               consistency_checker.py:119-212 */
  TVC_S_CC_THRESHOLD = 19,   /* stateless adaptive threshold                   consistency_checker.py:214-242 */
  TVC_S_CC_CONFIDENCE = 20,  /*                                                consistency_checker.py:244-272 */
  TVC_S_N_RET = 21,          /* retrieval references used (after de-duplication) */
  TVC_S_N_GEN = 22,
  TVC_S_REF_SIGMA = 23       /* std of cos(image, all references): the README sigma-rule, README.md:474-482 */
};

/* decision bits written per query by tvc_consistency_* */
#define TVC_FLAG_DET_ADV 1u   /* aggregated > detection_threshold      detector.py:399 */
#define TVC_FLAG_CC_ADV 2u    /* overall < threshold                   consistency_checker.py:93 */
#define TVC_FLAG_SIGMA_ADV 4u /* ref sigma > sigma_threshold           README.md:846 */

typedef struct {
  int32_t n_variants;          /* V, stride of the variant arrays (<= TVC_MAX_VARIANTS) */
  int32_t n_retrieval;         /* R, stride of the retrieval arrays (<= TVC_MAX_REFS) */
  int32_t n_generative;        /* G, stride of the generative arrays (<= TVC_MAX_REFS) */
  uint32_t methods;            /* bit0 text_variants, bit1 sd_reference, bit2 consistency (DetectorConfig.detection_methods) */
  int32_t aggregation;         /* 0 weighted_mean, 1 mean, 2 max, 3 min (DetectorConfig.score_aggregation) */
  float w_text_variants;       /* 0.4  detector.py:662-666 */
  float w_sd_reference;        /* 0.4 */
  float w_consistency;         /* 0.2 */
  float detection_threshold;   /* 0.5  detector.py:194 */
  int32_t voting;              /* 0 simple, 1 weighted, 2 adaptive (ConsistencyChecker.voting_strategy) */
  float cc_weights[4];         /* original, text_variant, retrieval, generative: 0.25 each */
  float cc_base_threshold;     /* 0.5 */
  int32_t cc_adaptive;         /* 1: apply the stateless part of _get_adaptive_threshold */
  float dedup_threshold;       /* 0.95 defenses/detector.py:318; <= -1 disables feature de-duplication */
  float sigma_threshold;       /* 0.30 README.md:846 */
} tvc_detector_params;

int tvc_version(void);
const char* tvc_status_string(int status);
void tvc_detector_params_default(tvc_detector_params* p);

int tvc_ctx_create(int device, tvc_ctx** out);
int tvc_ctx_destroy(tvc_ctx* ctx);
const char* tvc_last_error(tvc_ctx* ctx);
/* kernels launched through this context since creation (bench.py's gpu_launches) */
int64_t tvc_ctx_launch_count(tvc_ctx* ctx);
/* tuning knobs: "pair_min_rows" = query rows from which tvc_search uses the CTA-pair (cta_group::2)
 * kernel instead of the single-CTA one (default 4096; 0 = always, INT64_MAX = never);
 * "pace_every" / "pace_ahead" = the CTA pairs of a wave of the pair kernel stay within pace_ahead blocks
 * of pace_every gallery tiles of each other, so that a gallery tile is read from HBM once per wave
 * (defaults 8 / 2; pace_every = 0 switches pacing off; also TVC_PACE_EVERY / TVC_PACE_AHEAD in the
 * environment at context creation);
 * "rq_min_tiles" = units of at least this many 256-row gallery tiles run on the pair kernel that keeps half of
 * the query tile resident in shared memory (default 64; INT64_MAX = never; TVC_RQ_MIN_TILES), "rq_resident" =
 * resident k-blocks 5..7 (0 = by dimension); "ts_min_tiles" = the same for the revision that keeps the query
 * tile in tensor memory (measured slower: default never; TVC_TS_MIN_TILES); "debug_flags" = profiling only;
 * "kocc_part_min" = tvc_k_occurrence takes its bucketed two-pass path (bucket-sort into 16-bit keys, count with
 * shared-memory atomics) for index streams of at least this many entries (default 1 Mi; 0 = whenever the path
 * applies: 16-byte aligned stream, histogram beyond shared memory and up to 4 161 536 bins; INT64_MAX = never;
 * TVC_KOCC_PART_MIN).  It borrows 2 bytes per entry of the stream's workspace.
 * Results do not depend on any of them (every kernel revision is bit-identical to the others). */
int tvc_ctx_set_option(tvc_ctx* ctx, const char* name, int64_t value);
/* Frees the context's grow-only per-stream device workspaces (query operands, candidate lists, host
 * staging windows) after synchronising the device; they are re-grown on demand.  For callers that
 * alternate between very different batch sizes and want the HBM back in between (new: the
 * reference holds no device memory of its own; the Python mirrors call it from clear_cache(), the
 * hook src/pipeline.py:742-780 already invokes).  `freed_bytes` may be NULL. */
int tvc_ctx_release_workspace(tvc_ctx* ctx, int64_t* freed_bytes);
/* CUDA-event time in ms of the last gemm_topk launch made with timing enabled (roofline.achieved) */
int tvc_ctx_set_timing(tvc_ctx* ctx, int enabled);
int tvc_ctx_last_search_kernel_ms(tvc_ctx* ctx, float* ms, int64_t* launches);

/* Gallery: N rows of dimension d resident in HBM as bf16 [N, d_pad] (GEMM operand, d_pad = d
 * rounded up to 64) plus an fp32 master [N, d] used to re-rank the bf16 candidates.
 * `global_row_offset` is added to every returned index (row-sharded galleries).
 * Replaces: faiss.IndexFlatIP(d) + index.add(features)  src/retrieval.py:477-525 (FaissIndexManager
 * .build_index/.add_to_index :117-155, RetrievalIndex.build_index/.add_items :212-287,
 * experiments/defenses/retrieval_ref.py:140-156); the per-query np.array([ref.vector ...]) of
 * ReferenceBank  src/ref_bank.py:475; references.append / pop  src/ref_bank.py:143-151, 365-399. */
int tvc_gallery_create(tvc_ctx* ctx, const void* rows, int dtype, int64_t n, int32_t d,
                       int64_t global_row_offset, uint32_t flags, int64_t capacity_hint, void* stream,
                       tvc_gallery** out);
int tvc_gallery_append(tvc_gallery* g, const void* rows, int dtype, int64_t n, void* stream);
int tvc_gallery_truncate(tvc_gallery* g, int64_t n);
/* copy row `src` over row `dst` (swap-with-last removal used by the ReferenceBank eviction policies,
 * src/ref_bank.py:365-399) */
int tvc_gallery_move_row(tvc_gallery* g, int64_t src, int64_t dst, void* stream);
int tvc_gallery_info(const tvc_gallery* g, int64_t* n, int32_t* d, int64_t* global_row_offset,
                     uint32_t* flags);
/* device pointers of the resident copies (for zero-copy consumers; may be NULL when absent) */
int tvc_gallery_device_ptrs(const tvc_gallery* g, const void** bf16_rows, int32_t* d_pad,
                            const float** f32_rows);
/* gather rows (local indices) as fp32 [n, d] into host or device memory */
int tvc_gallery_get_rows(tvc_gallery* g, const int64_t* idx, int64_t n, float* out, void* stream);
int tvc_gallery_destroy(tvc_gallery* g);
/* Non-owning view over caller-owned fp32 device rows [n, d] (rows fetched from peer shards): usable
 * as ret_gallery / gen_gallery of tvc_consistency_emb and with tvc_gallery_get_rows; not searchable. */
int tvc_gallery_wrap_f32(tvc_ctx* ctx, const float* device_rows, int64_t n, int32_t d,
                         int64_t global_row_offset, tvc_gallery** out);

/* Row-sharded galleries across the GPUs of one box (one process per GPU).  A rank exports its fp32
 * master as a CUDA IPC handle, peers import it as a view, and a GROUP of [own shard + peer views]
 * is passed as ret_gallery / gen_gallery of tvc_consistency_emb, whose kernel then reads the
 * referenced rows of other shards directly from peer HBM over NVLink (no staging collective).
 * Exchange the 64-byte handles with any host channel (torch.distributed.all_gather_object). */
#define TVC_IPC_HANDLE_BYTES 64
int tvc_gallery_export_ipc(tvc_gallery* g, void* handle /* TVC_IPC_HANDLE_BYTES */);
int tvc_gallery_import_ipc(tvc_ctx* ctx, const void* handle, int64_t n, int32_t d, int64_t global_row_offset,
                           tvc_gallery** out);
/* group of up to 16 plain galleries / views with disjoint global index ranges; owns nothing */
int tvc_gallery_group_create(tvc_ctx* ctx, tvc_gallery** parts, int32_t n_parts, tvc_gallery** out);

/* Top-k of every query row against the gallery (IndexFlatIP semantics): out_sim [m, k] f32, out_idx [m, k] i64.
 * Entries with similarity < threshold are dropped (pass -INFINITY for none). 1 <= k <= TVC_MAX_K (56; larger k:
 * TVC_ERR_UNSUPPORTED - the Python mirrors then take tvc_similarity_matrix + a chunked top-k).
 * How exact: a row's candidates are its KP best gallery rows by the bf16 tensor-core score (KP = 16 / 32 / 64 for
 * k <= 10 / 26 / 56, per gallery range); they are re-scored from the fp32 masters, so the returned similarities and
 * the order among the candidates are fp32.  A true top-k row can be missed only if bf16 rounding (~1e-3 for unit
 * rows) pushes it below bf16 rank KP, i.e. only among rows within ~1e-3 of the k-th similarity - the band in which
 * north_star allows index disagreement.  Galleries built with TVC_GALLERY_NO_MASTER return the bf16 scores.
 * Replaces: index.search(q, k)  src/retrieval.py:652-656 (:130-137, :235-264); sklearn
 * cosine_similarity + argsort  src/retrieval.py:669-671; ReferenceBank._compute_similarities + `>= thr`
 * + argsort  src/ref_bank.py:462-484, 191-203; _faiss_retrieve / _numpy_retrieve
 * experiments/defenses/retrieval_ref.py:246-290; F.cosine_similarity + topk(.,1)
 * src/attacks/hubness_attack.py:482-489; the k-NN of the hubness spec
 * references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:43-47 (TVC_SEARCH_SKIP_SELF). */
int tvc_search(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m, int32_t d,
               int32_t k, float threshold, uint32_t flags, float* out_sim, int64_t* out_idx,
               void* stream);

/* ---- Sharded exact search (gallery row-sharded over the GPUs of one box, one process per GPU) ----
 * Phase 1, every rank, all query rows: GEMM + per-range top-KP on this rank's shard; the KP =
 * tvc_candidate_width(k) best candidates of each row by GEMM score (global indices, -1 = unused) go
 * either to the local lists cand_val/cand_idx [m, KP] or - with `scatter` - straight into the receive
 * buffers of the ranks that own the query slices: row r belongs to slice j = r / rows_per_slice and is
 * written at val[j][(slot * rows_in_slice_j + r - j*rows_per_slice) * KP ...].  val[j] / idx[j] are
 * device pointers of THIS process: the rank's own buffer or a peer's buffer opened with tvc_peer_open,
 * so the exchange is the kernel's own store stream over NVLink (no collective).  The caller orders
 * phase 2 after every rank's phase 1 (a barrier on the stream) and double-buffers the receive buffers.
 * Phase 2, the owner of a slice: merge the `parts` lists of each of its rows ([parts, m, KP]), re-score
 * the KP best in fp32 from the masters behind `g` (a gallery or a group with peer views) and emit the
 * top-k (score desc, index asc, `threshold` filter) - bit-identical to tvc_search on the unsharded
 * gallery.  Device pointers only. */
/* Query operand of the GEMM, prepared once per batch and BROADCAST over peer memory: the rows
 * (device, any dtype, optionally L2-normalised) are converted to the bf16 zero-padded row format
 * (tvc_query_row_bytes(d) bytes per row) and written at row `dst_row0` of each of the n_dst buffers -
 * the rank's own query buffer and its peers' (tvc_peer_alloc / tvc_peer_open).  A rank that holds only
 * its slice of the batch thereby hands the operand to the whole box; pass the filled buffer to
 * tvc_search_candidates with TVC_SEARCH_PREPARED_Q (after a barrier when peers wrote into it). */
int tvc_query_row_bytes(int32_t d);
int tvc_prepare_queries(tvc_ctx* ctx, const void* rows, int dtype, int64_t m, int32_t d, uint32_t flags,
                        int32_t n_dst, void* const* dst, int64_t dst_row0, void* stream);

typedef struct {
  int32_t n_slices;
  int32_t slot;
  int64_t rows_per_slice;
  float* val[16];
  int64_t* idx[16];
} tvc_scatter;
int tvc_candidate_width(int32_t k);
int tvc_search_candidates(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m, int32_t d,
                          int32_t k, uint32_t flags, const tvc_scatter* scatter, float* cand_val,
                          int64_t* cand_idx, void* stream);
int tvc_rerank_candidates(tvc_ctx* ctx, tvc_gallery* g, const float* queries, int64_t m, int32_t d,
                          int32_t parts, int32_t kp, const float* cand_val, const int64_t* cand_idx, int32_t k,
                          float threshold, float* out_sim, int64_t* out_idx, void* stream);
/* Phase 2 done where the rows live (instead of tvc_rerank_candidates pulling KP fp32 master rows per query row
 * over NVLink): three steps with a stream-ordered barrier between them, all buffers allocated with
 * tvc_peer_alloc and mapped by every rank.
 *  tvc_exchange_merge    (owner of a query slice, its m rows): the `parts` received lists [parts, m, KP] ->
 *                        the KP best by GEMM score; the index list of row r is stored at req_dst[s][r*KP ..]
 *                        for each of the n_dst shards (req_dst[s] = this owner's block [m, KP] inside shard
 *                        s's request area, own or peer memory).
 *  tvc_exchange_rescore  (every shard, all m_total rows of all owners): `req` is this shard's request area
 *                        [owners, rows_per_slice, KP]; entries that fall into `shard`'s row range are scored
 *                        in fp32 = <q_f32[row], master row> (q_f32 [m_total, d]: the fp32 query rows of the whole
 *                        batch, copied to every rank with tvc_peer_copy under the GEMM) and stored at
 *                        score_dst[owner][row_in_slice*KP + slot] (own or peer memory).
 *  tvc_exchange_finalize (owner): req [m, KP] (its own block) + score [m, KP] -> top-k ordered (score desc,
 *                        index asc), entries below `threshold` dropped - bit-identical to tvc_search on the
 *                        unsharded gallery.
 * Device pointers only.  New: the reference has no multi-GPU retrieval (src/retrieval.py:112,508). */
int tvc_exchange_merge(tvc_ctx* ctx, int64_t m, int32_t parts, int32_t kp, const float* cand_val,
                       const int64_t* cand_idx, int32_t n_dst, int64_t* const* req_dst, void* stream);
int tvc_exchange_rescore(tvc_ctx* ctx, tvc_gallery* shard, const float* q_f32, int32_t d, int32_t owners,
                         int64_t rows_per_slice, int64_t m_total, int32_t kp, const int64_t* req,
                         float* const* score_dst, void* stream);
int tvc_exchange_finalize(tvc_ctx* ctx, int64_t m, int32_t kp, int32_t k, float threshold, const int64_t* req,
                          const float* score, float* out_sim, int64_t* out_idx, void* stream);
/* asynchronous copy between device buffers of this rank and / or peer-mapped buffers (copy engines over
 * NVLink: no SM is taken from a kernel running beside it) */
int tvc_peer_copy(tvc_ctx* ctx, void* dst, const void* src, int64_t bytes, void* stream);
/* device buffers shareable with the other ranks of the box (cudaMalloc + CUDA IPC; zero-filled) */
int tvc_peer_alloc(tvc_ctx* ctx, int64_t bytes, void** ptr, void* handle /* TVC_IPC_HANDLE_BYTES */);
int tvc_peer_open(tvc_ctx* ctx, const void* handle, void** ptr);
int tvc_peer_close(tvc_ctx* ctx, void* ptr);
int tvc_peer_free(tvc_ctx* ctx, void* ptr);

/* Dense [m, N] fp32 similarity matrix (tcgen05 GEMM, plain store epilogue). Only for sizes that
 * fit; the search path never materialises it.
 * Replaces: MultiModalRetriever.compute_similarity_matrix  src/retrieval.py:682-722;
 * SimilarityCalculator.batch_cosine_similarity  src/utils/metrics.py:144-164. */
int tvc_similarity_matrix(tvc_ctx* ctx, tvc_gallery* g, const void* queries, int q_dtype, int64_t m,
                          int32_t d, uint32_t flags, float* out, void* stream);

/* (new: the reference has no multi-GPU retrieval, src/retrieval.py:112,508 are single-device)
 * Merge `parts` candidate lists per row (the all-gathered per-shard top-k): in_sim/in_idx
 * [m, parts, k] -> out [m, k], ordered (sim desc, idx asc); idx < 0 entries are ignored. */
int tvc_merge_topk(tvc_ctx* ctx, const float* in_sim, const int64_t* in_idx, int64_t m,
                   int32_t parts, int32_t k, float* out_sim, int64_t* out_idx, void* stream);

/* Variant-consistency reduction fed precomputed similarities.
 * Replaces (per query, all three decision stacks in one pass): AdversarialDetector
 * ._detect_by_text_variants / _sd_reference / _consistency, _aggregate_scores and the `> threshold`
 * decision  src/detector.py:441-590, 643-682, 399; ConsistencyChecker.make_decision (voting, stateless
 * adaptive threshold, `<`, confidence)  experiments/defenses/consistency_checker.py:74-272;
 * _compute_cross_modal_variance  experiments/defenses/detector.py:295-300.
 *  s0 [Q]; sv [Q,V]; sr [Q,R] with r_cnt [Q] (NULL = all R valid); sg [Q,G] with g_cnt [Q];
 *  sxv [Q, V*(V-1)/2] variant<->variant cosines (may be NULL).  scores [Q, TVC_NSCORES]; flags [Q]. */
int tvc_consistency_sims(tvc_ctx* ctx, const tvc_detector_params* p, int64_t q, const float* s0,
                         const float* sv, const float* sr, const int32_t* r_cnt, const float* sg,
                         const int32_t* g_cnt, const float* sxv, float* scores, uint8_t* flags,
                         void* stream);

/* Variant-consistency reduction fed embedding rows (fp32, L2-normalised by the encoders).
 * Replaces: MultiModalDefenseDetector._compute_consistency_scores and _deduplicate_references
 * experiments/defenses/detector.py:228-325 (plus everything tvc_consistency_sims replaces).
 *  img [Q,d]; txt [Q,d]; var [Q,V,d];
 *  retrieval refs: ret_idx [Q, n_ret_cand] global indices into `ret_gallery` (the search output,
 *  variant-major), greedily de-duplicated (index, then cosine > dedup_threshold) and cut to R;
 *  generative refs: either gen [Q,G,d] with g_cnt [Q], or gen_idx [Q, n_gen_cand] into `gen_gallery`.
 *  Optional outputs: out_sv [Q,V], out_sr [Q,R], out_sg [Q,G] (unused slots = 0). */
int tvc_consistency_emb(tvc_ctx* ctx, const tvc_detector_params* p, int64_t q, int32_t d,
                        const float* img, const float* txt, const float* var,
                        tvc_gallery* ret_gallery, const int64_t* ret_idx, int32_t n_ret_cand,
                        const float* gen, const int32_t* g_cnt, tvc_gallery* gen_gallery,
                        const int64_t* gen_idx, int32_t n_gen_cand, float* scores, uint8_t* flags,
                        float* out_sv, float* out_sr, float* out_sg, void* stream);

/* The documented TVC reference-vector rule (README.md:474-482, 846; prose only in the reference): for
 * every text variant v the rows ret_idx [Q, V, k] of `ret_gallery` (-1 = unused) and the generated rows
 * gen [Q, V, m, d] are averaged into a per-variant reference vector r_v, the r_v into the Reference
 * Vector r;  out_s [Q, V] = cos(image, r_v) (0 for a variant without rows), out_ref [Q] = cos(image, r)
 * (may be NULL), out_sigma [Q] = std_v(out_s) (population, fp64), flags [Q] = TVC_FLAG_SIGMA_ADV iff
 * sigma > sigma_threshold.  k + m <= 32, V <= TVC_MAX_VARIANTS; fp32 masters required. */
int tvc_reference_vector_rule(tvc_ctx* ctx, int64_t q, int32_t d, int32_t v, const float* img,
                              tvc_gallery* ret_gallery, const int64_t* ret_idx, int32_t k, const float* gen,
                              int32_t m, float sigma_threshold, float* out_s, float* out_ref, float* out_sigma,
                              uint8_t* flags, void* stream);

/* Retrieval quality of ranked lists (src/utils/metrics.py:386-574, binary relevance).  topk_idx [q, k]
 * is the search output (rank order, -1 = unused); the relevant items of query i are
 * rel_idx[rel_ptr[i] .. rel_ptr[i+1]).  out [q, 2 + 3*n_k] per query: reciprocal rank, average
 * precision, then recall@K, precision@K, NDCG@K for each of the n_k <= 8 values K <= k.  RR and AP
 * are those of the full ranking whenever every relevant item lies inside the top-k list (AP is
 * normalised by the number of relevant items; a first hit beyond k gives RR = 0). */
int tvc_retrieval_metrics(tvc_ctx* ctx, const int64_t* topk_idx, int64_t q, int32_t k, const int64_t* rel_ptr,
                          const int64_t* rel_idx, int64_t n_rel, const int32_t* k_values, int32_t n_k, float* out,
                          void* stream);

/* k-occurrence histogram N_k(j) = #{rows i : j in idx[i, :k]}; idx < 0 or >= n_bins ignored.
 * Replaces: the `hubness_counts[j] += 1` double loop
 * references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:48-57; `(top1 == 0).sum()`
 * src/attacks/hubness_attack.py:492-496; the per-image scores of
 * benchmarks/hubness_attack_benchmark.py:335-348.
 * counts [n_bins] int32; zero_first != 0 clears it before accumulating. `idx_base` is subtracted
 * from every index first (a rank histogramming global indices into its own slice passes 0). */
int tvc_k_occurrence(tvc_ctx* ctx, const int64_t* idx, int64_t m, int32_t k, int64_t idx_base,
                     int64_t n_bins, int32_t* counts, int zero_first, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TVC_H_ */
