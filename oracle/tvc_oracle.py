"""CPU oracle for the TVC scoring + retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline — never as (part of) the product
path.  The product path is libtvc.so (CUDA, sm_100a) and fails loudly without it.

This is a restatement in NumPy (fp32 arithmetic with fp64 statistics, as the reference does it) of
the reference's algorithm for the path; every function cites the reference file:line it follows
(paths relative to the reference tree).  The reference is 100 % Python and ships no tests, golden
vectors or known-answer fixtures for this path, and the arithmetic at its
boundary lives in third-party libraries that are not vendored:
  * FAISS  (README.md:546 pins faiss-gpu==1.7.4; not installable here)  IndexFlatIP.search
    = exact inner-product top-k, published algorithm: sgemm + per-row heap, descending order,
    -1 / -inf padding when fewer than k rows exist.
  * scikit-learn >=1.3 cosine_similarity (requirements.txt:21), SciPy >=1.10 cosine
    (requirements.txt:20), torch mm/topk/cosine_similarity.
Parity is pinned by running the reference's OWN classes in this container (ReferenceBank,
ConsistencyChecker, SimilarityCalculator import standalone; retrieval / detector / hubness import
with shim modules) and committing their outputs as fixtures: tests/golden/make_golden.py ->
tests/golden/*.npz, checked by tests/test_oracle_golden.py.

Tie rule (BASELINE.json north_star): order is (similarity descending, index ascending).  The
reference's numpy fallbacks (`np.argsort(s)[::-1]`, src/retrieval.py:670, src/ref_bank.py:202) do
not define a tie order; FAISS's heap does not either.  Everything here is explicit about it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NSCORES = 24
# column indices of the score matrix (mirror include/tvc.h)
(S_ORIGINAL, S_TV_MEAN, S_TV_STD, S_TV_MIN, S_TV_VAR, S_RET_MEAN, S_RET_STD, S_GEN_MEAN, S_GEN_STD,
 S_GEN_MAX, S_CROSS_MODAL_VAR, S_XV_MEAN, S_XV_MIN, S_XV_VAR, S_DET_TV, S_DET_SD, S_DET_C, S_DET_AGG,
 S_CC_OVERALL, S_CC_THRESHOLD, S_CC_CONFIDENCE, S_N_RET, S_N_GEN, S_REF_SIGMA) = range(NSCORES)
FLAG_DET_ADV, FLAG_CC_ADV, FLAG_SIGMA_ADV = 1, 2, 4


# --------------------------------------------------------------------------------------------
# similarity + top-k  (kernel a)
# --------------------------------------------------------------------------------------------
def l2_normalize(x: np.ndarray) -> np.ndarray:
    """Row L2 normalisation; zero rows stay zero (sklearn.preprocessing.normalize semantics used by
    cosine_similarity, src/retrieval.py:669,706)."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    inv = np.where(n > 0, 1.0 / np.where(n > 0, n, 1.0), 0.0)
    return (x * inv).astype(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even) and return as fp32 — the GEMM operand precision."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def similarity_matrix(q: np.ndarray, g: np.ndarray, metric: str = "dot_product") -> np.ndarray:
    """src/retrieval.py:682-722 compute_similarity_matrix; src/utils/metrics.py:144-164."""
    q = np.asarray(q, dtype=np.float32)
    g = np.asarray(g, dtype=np.float32)
    if metric == "cosine":
        q, g = l2_normalize(q), l2_normalize(g)
    elif metric == "euclidean":
        d2 = (q.astype(np.float64) ** 2).sum(1)[:, None] + (g.astype(np.float64) ** 2).sum(1)[None, :] \
            - 2.0 * (q.astype(np.float64) @ g.astype(np.float64).T)
        return (1.0 / (1.0 + np.sqrt(np.maximum(d2, 0.0)))).astype(np.float32)
    return q @ g.T


def topk_rows(sims: np.ndarray, k: int, threshold: float = -np.inf,
              index_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row top-k ordered (similarity desc, index asc); unused slots (-inf, -1).

    Follows IndexFlatIP.search as called at src/retrieval.py:652-656 (returns k columns even when
    the gallery has fewer rows, padding with -1) and the `>= threshold` filter of
    src/ref_bank.py:197.
    """
    sims = np.asarray(sims, dtype=np.float32)
    m, n = sims.shape
    out_s = np.full((m, k), -np.inf, dtype=np.float32)
    out_i = np.full((m, k), -1, dtype=np.int64)
    if n == 0:
        return out_s, out_i
    kk = min(k, n)
    idx = np.arange(n, dtype=np.int64)
    for r in range(m):
        row = sims[r]
        if kk < n:
            # candidates: everything >= the kk-th largest value (keeps all ties), then exact order
            kth = np.partition(row, n - kk)[n - kk]
            cand = idx[row >= kth]
        else:
            cand = idx
        order = np.lexsort((cand, -row[cand].astype(np.float64)))
        sel = cand[order][:kk]
        vals = row[sel]
        keep = vals >= threshold
        cnt = int(keep.sum())
        # values are sorted descending, so the kept ones are a prefix
        out_s[r, :cnt] = vals[:cnt]
        out_i[r, :cnt] = sel[:cnt] + index_offset
    return out_s, out_i


def search(q: np.ndarray, g: np.ndarray, k: int, metric: str = "dot_product",
           threshold: float = -np.inf, index_offset: int = 0, skip_self: bool = False,
           chunk: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k of every query row over the gallery, fp32 (`index.search`,
    src/retrieval.py:652-656; sklearn fallback :669-671; retrieval_ref.py:246-290).

    skip_self: query set == gallery, drop column == row (the `[:, 1:k+1]` of the hubness spec,
    references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:46)."""
    q = np.asarray(q, dtype=np.float32)
    g = np.asarray(g, dtype=np.float32)
    if metric == "cosine":
        q, g = l2_normalize(q), l2_normalize(g)
    m = q.shape[0]
    out_s = np.full((m, k), -np.inf, dtype=np.float32)
    out_i = np.full((m, k), -1, dtype=np.int64)
    for r0 in range(0, m, chunk):
        s = q[r0:r0 + chunk] @ g.T
        if skip_self:
            rows = np.arange(s.shape[0])
            cols = rows + r0 - index_offset
            ok = (cols >= 0) & (cols < g.shape[0])
            s[rows[ok], cols[ok]] = -np.inf
        cs, ci = topk_rows(s, k, threshold, index_offset)
        if skip_self:
            dead = ~np.isfinite(cs)
            ci[dead] = -1
        out_s[r0:r0 + chunk], out_i[r0:r0 + chunk] = cs, ci
    return out_s, out_i


def ref_bank_similarities(ref_vectors: np.ndarray, query: np.ndarray) -> np.ndarray:
    """src/ref_bank.py:462-484 _compute_similarities: dot / (|r| |q| + 1e-8), in the dtype given."""
    ref_vectors = np.asarray(ref_vectors)
    query = np.asarray(query)
    qn = np.linalg.norm(query)
    rn = np.linalg.norm(ref_vectors, axis=1)
    return np.dot(ref_vectors, query) / (rn * qn + 1e-8)


def ref_bank_query(ref_vectors: np.ndarray, query: np.ndarray, top_k: int = 10,
                   similarity_threshold: Optional[float] = None,
                   config_threshold: float = 0.9) -> Tuple[np.ndarray, np.ndarray]:
    """src/ref_bank.py:172-224 query_similar: `threshold = similarity_threshold or config` (an
    explicit 0.0 falls back to the config value, :191), keep >= threshold, sort descending, take k.
    Returns (indices, similarities); ties ordered by lower index."""
    if len(ref_vectors) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    thr = similarity_threshold or config_threshold
    s = ref_bank_similarities(ref_vectors, query)
    valid = np.where(s >= thr)[0]
    if len(valid) == 0:
        return np.zeros(0, np.int64), np.zeros(0, s.dtype)
    order = np.lexsort((valid, -s[valid]))
    top = valid[order][:top_k]
    return top.astype(np.int64), s[top]


def merge_topk(sims: np.ndarray, idx: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge [m, parts, k] per-shard candidates into the global top-k (new: SURVEY.md §8e)."""
    m = sims.shape[0]
    s = sims.reshape(m, -1)
    i = idx.reshape(m, -1)
    out_s = np.full((m, k), -np.inf, dtype=np.float32)
    out_i = np.full((m, k), -1, dtype=np.int64)
    for r in range(m):
        ok = i[r] >= 0
        ss, ii = s[r][ok], i[r][ok]
        # a global index appears once
        order = np.lexsort((ii, -ss.astype(np.float64)))
        ss, ii = ss[order], ii[order]
        if len(ii) > 1:
            first = np.ones(len(ii), dtype=bool)
            first[1:] = (ii[1:] != ii[:-1]) | (ss[1:] != ss[:-1])
            ss, ii = ss[first], ii[first]
        n = min(k, len(ii))
        out_s[r, :n], out_i[r, :n] = ss[:n], ii[:n]
    return out_s, out_i


# --------------------------------------------------------------------------------------------
# hubness  (kernel c)
# --------------------------------------------------------------------------------------------
def k_occurrence(idx: np.ndarray, n_bins: int, idx_base: int = 0) -> np.ndarray:
    """N_k(j) = #{queries i : j in topk(i)} — the double loop of
    references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:49-53 over precomputed top-k."""
    flat = np.asarray(idx, dtype=np.int64).ravel() - idx_base
    flat = flat[(flat >= 0) & (flat < n_bins)]
    return np.bincount(flat, minlength=n_bins).astype(np.int32)


def hubness_spec(features: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:30-58 `compute_hubness`:
    cosine matrix, k nearest neighbours excluding self, counts, counts / (N k)."""
    f = l2_normalize(features)
    _, idx = search(f, f, k, skip_self=True)
    counts = k_occurrence(idx, f.shape[0])
    return counts, counts.astype(np.float64) / (f.shape[0] * k)


def hubness_top1_fraction(image_features: np.ndarray, text_features: np.ndarray,
                          target: int = 0) -> float:
    """src/attacks/hubness_attack.py:464-498: fraction of texts whose top-1 image is `target`
    (the `k` argument is ignored there, :489)."""
    _, idx = search(text_features, image_features, 1, metric="cosine")
    return float((idx[:, 0] == target).sum()) / text_features.shape[0]


# --------------------------------------------------------------------------------------------
# variant-consistency reduction + decisions  (kernel b)
# --------------------------------------------------------------------------------------------
DEFAULT_PARAMS: Dict[str, object] = dict(
    n_variants=5, n_retrieval=10, n_generative=3, methods=7, aggregation=0,
    w_text_variants=0.4, w_sd_reference=0.4, w_consistency=0.2, detection_threshold=0.5,
    voting=1, cc_weights=(0.25, 0.25, 0.25, 0.25), cc_base_threshold=0.5, cc_adaptive=1,
    dedup_threshold=0.95, sigma_threshold=0.30)


def _f32(x: float) -> float:
    return float(np.float32(x))


def consistency_from_scores(s0, tv_c, tv_s, rt_c, rt_s, gn_c, gn_s, cmv, params=None,
                            threshold_history: Optional[Sequence[float]] = None):
    """experiments/defenses/consistency_checker.py:74-272 make_decision from the score dict:
    voting (:119-212), adaptive threshold (:214-242; the history blend :234-239 only when
    `threshold_history` is given — it is host-side state), decision (:93), confidence (:244-272).
    Returns (overall_score, threshold, confidence, is_adversarial)."""
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    four = [float(s0), float(tv_c), float(rt_c), float(gn_c)]
    valid = [s for s in four if s > 0]
    if p["voting"] == 0:
        overall = float(np.mean(valid)) if valid else 0.0
    else:
        if p["voting"] == 1:
            w = [float(np.float32(x)) for x in p["cc_weights"]]
        else:
            w = [1.0, 1.0 / (1.0 + tv_s), 1.0 / (1.0 + rt_s), 1.0 / (1.0 + gn_s)]
            t = sum(w)
            w = [x / t for x in w] if t > 0 else w
        ws = sum(s * wi for s, wi in zip(four, w) if s > 0)
        tw = sum(wi for s, wi in zip(four, w) if s > 0)
        overall = ws / tw if tw != 0 else 0.0
    thr = _f32(p["cc_base_threshold"])
    if p["cc_adaptive"]:
        if cmv > 0.1:
            thr += 0.1
        if (tv_s + rt_s + gn_s) / 3.0 > 0.2:
            thr += 0.05
        if threshold_history is not None and len(threshold_history) > 10:
            thr = 0.7 * thr + 0.3 * float(np.mean(list(threshold_history)[-10:]))
        thr = min(max(thr, 0.1), 0.9)
    cc_adv = overall < thr
    dist_conf = abs(overall - thr) / thr
    cons_conf = 1.0 - float(np.std(valid)) if len(valid) > 1 else 0.5
    var_conf = 1.0 - min(float(cmv), 1.0)
    conf = min(max((dist_conf + cons_conf + var_conf) / 3.0, 0.0), 1.0)
    return overall, thr, conf, cc_adv


def consistency_one(s0: float, sv: Sequence[float], sr: Sequence[float], sg: Sequence[float],
                    sxv: Sequence[float] = (), params: Optional[Dict[str, object]] = None
                    ) -> Tuple[np.ndarray, int]:
    """One query: all statistics and decisions, in float64 on the (fp32) similarities given.

    src/detector.py:441-501 (text variants), :503-557 (SD references), :559-590 (consistency),
    :643-682 (_aggregate_scores), :399 (decision);
    experiments/defenses/detector.py:228-300 (_compute_consistency_scores, cross-modal variance);
    experiments/defenses/consistency_checker.py:119-272 (voting, stateless adaptive threshold,
    confidence), :93 (decision); experiments/defenses/text_variants.py:412-451 (variant pairs);
    README.md:474-482,846 (sigma rule).
    """
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    s0 = float(s0)
    sv = np.asarray(sv, dtype=np.float64)
    sr = np.asarray(sr, dtype=np.float64)
    sg = np.asarray(sg, dtype=np.float64)
    sx = np.asarray(sxv, dtype=np.float64)
    out = np.zeros(NSCORES, dtype=np.float64)

    # --- AdversarialDetector
    det_tv = 0.0
    if len(sv):
        consistency = 1.0 - abs(s0 - sv.mean())
        variability = 1.0 - sv.std()
        det_tv = 1.0 - (consistency * 0.7 + variability * 0.3)
    det_sd = 1.0 - sg.mean() if len(sg) else 0.0
    det_c = 1.0 - s0
    scores = [det_tv, det_sd, det_c]
    weights = [p["w_text_variants"], p["w_sd_reference"], p["w_consistency"]]
    used = [i for i in range(3) if int(p["methods"]) & (1 << i)]
    agg = 0.0
    if used:
        vals = [scores[i] for i in used]
        if p["aggregation"] == 0:
            tw = sum(float(np.float32(weights[i])) for i in used)
            agg = sum(scores[i] * float(np.float32(weights[i])) for i in used) / tw if tw > 0 else 0.0
        elif p["aggregation"] == 2:
            agg = max(vals)
        elif p["aggregation"] == 3:
            agg = min(vals)
        else:
            agg = float(np.mean(vals))
    det_adv = agg > _f32(p["detection_threshold"])

    # --- MultiModalDefenseDetector scores
    tv_c = sv.mean() if len(sv) else s0
    tv_s = sv.std() if len(sv) else 0.0
    rt_c, rt_s = (sr.mean(), sr.std()) if len(sr) else (0.0, 0.0)
    gn_c, gn_s = (sg.mean(), sg.std()) if len(sg) else (0.0, 0.0)
    four = [s0, tv_c, rt_c, gn_c]
    valid = [s for s in four if s > 0]
    cmv = float(np.var(valid)) if len(valid) >= 2 else 0.0

    # --- ConsistencyChecker
    overall, thr, conf, cc_adv = consistency_from_scores(s0, tv_c, tv_s, rt_c, rt_s, gn_c, gn_s, cmv, p)

    refs = np.concatenate([sr, sg])
    sigma = float(refs.std()) if len(refs) else 0.0
    sig_adv = sigma > _f32(p["sigma_threshold"])

    out[S_ORIGINAL] = s0
    out[S_TV_MEAN], out[S_TV_STD] = tv_c, tv_s
    out[S_TV_MIN] = sv.min() if len(sv) else s0
    out[S_TV_VAR] = sv.var() if len(sv) else 0.0
    out[S_RET_MEAN], out[S_RET_STD] = rt_c, rt_s
    out[S_GEN_MEAN], out[S_GEN_STD] = gn_c, gn_s
    out[S_GEN_MAX] = sg.max() if len(sg) else 0.0
    out[S_CROSS_MODAL_VAR] = cmv
    if len(sx):
        out[S_XV_MEAN], out[S_XV_MIN], out[S_XV_VAR] = sx.mean(), sx.min(), sx.var()
    out[S_DET_TV], out[S_DET_SD], out[S_DET_C], out[S_DET_AGG] = det_tv, det_sd, det_c, agg
    out[S_CC_OVERALL], out[S_CC_THRESHOLD], out[S_CC_CONFIDENCE] = overall, thr, conf
    out[S_N_RET], out[S_N_GEN], out[S_REF_SIGMA] = len(sr), len(sg), sigma
    flag = (FLAG_DET_ADV if det_adv else 0) | (FLAG_CC_ADV if cc_adv else 0) | (FLAG_SIGMA_ADV if sig_adv else 0)
    return out, flag


def consistency_sims(s0, sv=None, sr=None, r_cnt=None, sg=None, g_cnt=None, sxv=None,
                     params: Optional[Dict[str, object]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Batched similarity-fed reduction: scores [Q, NSCORES] float64, flags [Q] uint8."""
    q = len(s0)
    scores = np.zeros((q, NSCORES), dtype=np.float64)
    flags = np.zeros(q, dtype=np.uint8)
    for i in range(q):
        v = sv[i] if sv is not None else ()
        r = sr[i][: (r_cnt[i] if r_cnt is not None else len(sr[i]))] if sr is not None else ()
        g = sg[i][: (g_cnt[i] if g_cnt is not None else len(sg[i]))] if sg is not None else ()
        x = sxv[i] if sxv is not None else ()
        scores[i], flags[i] = consistency_one(float(s0[i]), v, r, g, x, params)
    return scores, flags


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """torch.cosine_similarity(a, b, dim=-1): x.y / max(|x| |y|, 1e-8), fp32 in, fp32 out
    (experiments/defenses/detector.py:240,249,262,275)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    dot = (a.astype(np.float64) * b.astype(np.float64)).sum(-1)
    den = np.maximum(np.sqrt((a.astype(np.float64) ** 2).sum(-1)) * np.sqrt((b.astype(np.float64) ** 2).sum(-1)), 1e-8)
    return (dot / den).astype(np.float32)


def select_refs(img: np.ndarray, rows: np.ndarray, cand: Sequence[int], cap: int,
                dedup_threshold: float, offset: int = 0) -> Tuple[List[int], List[float]]:
    """Greedy reference selection, experiments/defenses/detector.py:184-204 + :302-325: walk the
    candidates in order, skip invalid / repeated indices and rows whose cosine to a kept row is
    > dedup_threshold, stop at `cap`.  Returns (kept local indices, cos(image, row))."""
    kept: List[int] = []
    sims: List[float] = []
    n = rows.shape[0]
    for c in cand:
        if len(kept) >= cap:
            break
        gi = int(c) - offset
        if gi < 0 or gi >= n or gi in kept:
            continue
        if dedup_threshold > -1.0 and kept:
            cs = cosine_rows(rows[kept], np.broadcast_to(rows[gi], (len(kept), rows.shape[1])))
            if bool((cs > np.float32(dedup_threshold)).any()):
                continue
        kept.append(gi)
        sims.append(float(cosine_rows(img[None, :], rows[gi][None, :])[0]))
    return kept, sims


def consistency_emb(img: np.ndarray, txt: np.ndarray, var: Optional[np.ndarray],
                    ret_rows: Optional[np.ndarray] = None, ret_idx: Optional[np.ndarray] = None,
                    gen: Optional[np.ndarray] = None, g_cnt: Optional[np.ndarray] = None,
                    gen_rows: Optional[np.ndarray] = None, gen_idx: Optional[np.ndarray] = None,
                    params: Optional[Dict[str, object]] = None, ret_offset: int = 0,
                    gen_offset: int = 0):
    """Batched embedding-fed reduction (the full `_compute_consistency_scores`,
    experiments/defenses/detector.py:228-293).  Returns scores, flags, (sv, sr, sg) lists."""
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    q = img.shape[0]
    scores = np.zeros((q, NSCORES), dtype=np.float64)
    flags = np.zeros(q, dtype=np.uint8)
    all_sv, all_sr, all_sg = [], [], []
    for i in range(q):
        s0 = float(cosine_rows(img[i][None], txt[i][None])[0])
        sv: List[float] = []
        sx: List[float] = []
        if var is not None and var.shape[1] > 0:
            v = var[i]
            sv = [float(x) for x in cosine_rows(np.broadcast_to(img[i], v.shape), v)]
            for a in range(v.shape[0]):
                for b in range(a + 1, v.shape[0]):
                    sx.append(float(cosine_rows(v[a][None], v[b][None])[0]))
        sr: List[float] = []
        if ret_rows is not None and ret_idx is not None:
            _, sr = select_refs(img[i], ret_rows, ret_idx[i], int(p["n_retrieval"]),
                                float(p["dedup_threshold"]), ret_offset)
        sg: List[float] = []
        if gen is not None:
            ng = int(g_cnt[i]) if g_cnt is not None else gen.shape[1]
            ng = max(0, min(ng, int(p["n_generative"])))
            sg = [float(x) for x in cosine_rows(np.broadcast_to(img[i], gen[i][:ng].shape), gen[i][:ng])]
        elif gen_rows is not None and gen_idx is not None:
            _, sg = select_refs(img[i], gen_rows, gen_idx[i], int(p["n_generative"]),
                                float(p["dedup_threshold"]), gen_offset)
        scores[i], flags[i] = consistency_one(s0, sv, sr, sg, sx, p)
        all_sv.append(sv)
        all_sr.append(sr)
        all_sg.append(sg)
    return scores, flags, (all_sv, all_sr, all_sg)


# --------------------------------------------------------------------------------------------
# small helpers of the path
# --------------------------------------------------------------------------------------------
def topk_overlap(indices1: np.ndarray, indices2: np.ndarray, k: int) -> float:
    """src/retrieval.py:178-183 compute_top_k_consistency."""
    return len(set(indices1[:k].tolist()) & set(indices2[:k].tolist())) / k


def similarity_distribution(s: np.ndarray) -> Dict[str, float]:
    """src/retrieval.py:164-172."""
    s = np.asarray(s)
    return dict(mean=float(np.mean(s)), std=float(np.std(s)), min=float(np.min(s)),
                max=float(np.max(s)), median=float(np.median(s)))


def scalar_cosine(x: np.ndarray, y: np.ndarray) -> float:
    """src/utils/metrics.py:116-141 SimilarityCalculator.cosine_similarity (zero vector -> 0.0)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    nx, ny = np.linalg.norm(x), np.linalg.norm(y)
    if nx == 0 or ny == 0:
        return 0.0
    return float(1.0 - (1.0 - np.dot(x, y) / (nx * ny)))


# --------------------------------------------------------------------------------------------
# synthetic workloads — shared by tests and bench.py so both sides see one input
# --------------------------------------------------------------------------------------------
def synth_gallery(n: int, d: int, seed: int = 42, clusters: int = 1024, noise: float = 0.35,
                  dup_rate: float = 1e-4) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = l2_normalize(rng.standard_normal((min(clusters, max(n, 1)), d), dtype=np.float32))
    assign = rng.integers(0, c.shape[0], size=n)
    g = c[assign] + (noise / math.sqrt(d)) * rng.standard_normal((n, d), dtype=np.float32)
    g = l2_normalize(g)
    ndup = int(n * dup_rate)
    if ndup and n > 2:
        src = rng.integers(0, n, size=ndup)
        dst = rng.integers(0, n, size=ndup)
        g[dst] = g[src]
    return g


def synth_queries(gallery: np.ndarray, q: int, v: int, seed: int = 123, q_noise: float = 0.5,
                  v_noise: float = 0.15):
    """Returns img [Q,d], txt [Q,d], var [Q,V,d] (all L2-normalised fp32)."""
    rng = np.random.default_rng(seed)
    n, d = gallery.shape
    pick = rng.integers(0, n, size=q)
    base = gallery[pick]
    sd = 1.0 / math.sqrt(d)
    txt = l2_normalize(base + q_noise * sd * rng.standard_normal((q, d), dtype=np.float32))
    img = l2_normalize(base + q_noise * sd * rng.standard_normal((q, d), dtype=np.float32))
    var = txt[:, None, :] + v_noise * sd * rng.standard_normal((q, v, d), dtype=np.float32)
    var = l2_normalize(var.reshape(q * v, d)).reshape(q, v, d)
    return img, txt, var


# --------------------------------------------------------------------------------------------
# Retrieval metrics from ranked lists (src/utils/metrics.py:386-574: binary relevance,
# DCG = sum rel_i / log2(i + 2), IDCG with all relevant items first, AP = mean precision at the
# ranks of the relevant items, RR = 1 / rank of the first relevant item).
def retrieval_metrics_from_topk(topk_idx: np.ndarray, relevant: Sequence[Sequence[int]],
                                k_values: Sequence[int]) -> np.ndarray:
    """Per query [rr, ap, recall@K.., precision@K.., ndcg@K..] with RR/AP restricted to the list
    (AP normalised by the number of relevant items)."""
    q, k = topk_idx.shape
    nk = len(k_values)
    out = np.zeros((q, 2 + 3 * nk), dtype=np.float64)
    for i in range(q):
        rel = set(int(x) for x in relevant[i])
        hits = np.array([int(j) >= 0 and int(j) in rel for j in topk_idx[i]], dtype=bool)
        pos = np.flatnonzero(hits)
        if len(pos):
            out[i, 0] = 1.0 / (pos[0] + 1)
            out[i, 1] = sum((n + 1) / (p + 1) for n, p in enumerate(pos)) / len(rel)
        for t, kv in enumerate(k_values):
            h = hits[:kv]
            nh = int(h.sum())
            out[i, 2 + t] = nh / len(rel) if rel else 0.0
            out[i, 2 + nk + t] = nh / kv
            dcg = sum(1.0 / math.log2(p + 2) for p in np.flatnonzero(h))
            idcg = sum(1.0 / math.log2(j + 2) for j in range(min(kv, len(rel))))
            out[i, 2 + 2 * nk + t] = dcg / idcg if idcg > 0 else 0.0
    return out


# --------------------------------------------------------------------------------------------
# The documented reference-vector rule (README.md:474-482, 846; prose only in the reference - parity for
# this per-variant form is against this restatement, UNPINNED.  The flat form of the same rule, std over
# cos(image, every reference), is consistency_one's S_REF_SIGMA and IS pinned: tests/golden/live_check.py
# exec's the pseudo-code of README.md:225-319 next to it):
#   per variant: mean of its top-k retrieved rows and its m generated rows -> r_v; mean_v r_v -> r;
#   S_v = cos(image, r_v); sigma = std(S) (population); adversarial iff sigma > threshold.
def reference_vector_rule(img: np.ndarray, ret_rows: Optional[np.ndarray], ret_idx: Optional[np.ndarray],
                          gen: Optional[np.ndarray], sigma_threshold: float = 0.30, ret_offset: int = 0):
    q = img.shape[0]
    v = ret_idx.shape[1] if ret_idx is not None else gen.shape[1]
    s = np.zeros((q, v), dtype=np.float64)
    ref = np.zeros(q, dtype=np.float64)
    sigma = np.zeros(q, dtype=np.float64)
    for i in range(q):
        means, valid = [], []
        for j in range(v):
            rows = []
            if ret_idx is not None:
                for gi in ret_idx[i, j]:
                    gi = int(gi) - ret_offset
                    if 0 <= gi < ret_rows.shape[0]:
                        rows.append(ret_rows[gi].astype(np.float32))
            if gen is not None:
                rows.extend(gen[i, j].astype(np.float32))
            if rows:
                r = np.mean(np.stack(rows).astype(np.float64), axis=0)
                means.append(r)
                valid.append(j)
                s[i, j] = float(np.dot(img[i].astype(np.float64), r) /
                                max(np.linalg.norm(img[i].astype(np.float64)) * np.linalg.norm(r), 1e-8))
        if valid:
            sigma[i] = float(np.std(s[i, valid]))
            rr = np.mean(np.stack(means), axis=0)
            ref[i] = float(np.dot(img[i].astype(np.float64), rr) /
                           max(np.linalg.norm(img[i].astype(np.float64)) * np.linalg.norm(rr), 1e-8))
    flags = (sigma > np.float32(sigma_threshold)).astype(np.uint8) * FLAG_SIGMA_ADV
    return s, ref, sigma, flags
