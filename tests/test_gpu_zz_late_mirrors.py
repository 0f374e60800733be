"""SimilarityCalculator mirror (src/utils/metrics.py:107-164) on the GPU: the batched cosine matrix is
tvc_similarity_matrix (bf16 operands, fp32 accumulate) - within 2e-3 of the fp32 reference arithmetic."""
import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu


def test_batch_cosine_similarity_matches_oracle(tvc_ctx):
    from multimodal_detection_consistency_b200.metrics import SimilarityCalculator
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((300, 512)) * 4).astype(np.float32)          # un-normalised, as callers pass them
    y = (rng.standard_normal((700, 512)) * 0.3).astype(np.float32)
    got = SimilarityCalculator.batch_cosine_similarity(x, y)
    want = O.similarity_matrix(x, y, "cosine")
    assert got.shape == (300, 700) and np.abs(got - want).max() <= 2e-3
    import torch
    got_t = SimilarityCalculator.batch_cosine_similarity(torch.from_numpy(x), torch.from_numpy(y).cuda())
    assert np.array_equal(got_t, got)
    assert abs(SimilarityCalculator.cosine_similarity(x[0], y[0]) - O.scalar_cosine(x[0], y[0])) <= 1e-6
    assert SimilarityCalculator.cosine_similarity(np.zeros(8), np.ones(8)) == 0.0


def test_retrieval_reference_generator_on_gpu(tvc_ctx, tmp_path):
    """defenses.RetrievalReferenceGenerator (experiments/defenses/retrieval_ref.py:34): features.npy database ->
    exact top-rerank_top_k search, similarity floor, cut to reference_count; the batch entry equals the single
    calls; statistics and cache bookkeeping as the reference's."""
    import json
    import torch
    from multimodal_detection_consistency_b200.defenses import RetrievalRefConfig, RetrievalReferenceGenerator
    rng = np.random.default_rng(23)
    n, d = 3000, 256
    feats = O.l2_normalize(rng.standard_normal((n, d), dtype=np.float32))
    np.save(tmp_path / "features.npy", feats)
    (tmp_path / "metadata.json").write_text(json.dumps([{"image_path": f"img_{j}.jpg"} for j in range(n)]))
    table = {f"text {j}": O.l2_normalize((feats[rng.integers(n)] + 0.05 * rng.standard_normal(d))[None].astype(np.float32))[0]
             for j in range(40)}
    clip = type("Clip", (), {"encode_text": lambda self, ts: torch.stack([torch.from_numpy(table[t]) for t in ts]) * 2.0})()
    gen = RetrievalReferenceGenerator(clip, str(tmp_path), RetrievalRefConfig(reference_count=5, similarity_threshold=0.3))
    names = list(table)
    singles = [gen.retrieve_references(t) for t in names[:10]]
    for t, refs in zip(names, singles):
        s, i = O.search(table[t][None], feats, 20, threshold=0.3)
        want = [int(j) for j in i[0][:5] if j >= 0]
        assert [r["index"] for r in refs] == want and len(refs) >= 1
        assert np.abs(np.array([r["similarity"] for r in refs]) - s[0][:len(refs)]).max() <= 2e-3
        assert refs[0]["metadata"] == {"image_path": f"img_{want[0]}.jpg"} and np.array_equal(refs[0]["features"], feats[want[0]])
    batch = gen.batch_retrieve_references(names)                      # 10 cache hits + 30 fresh in one launch
    assert [[r["index"] for r in refs] for refs in batch[:10]] == [[r["index"] for r in refs] for refs in singles]
    fresh = RetrievalReferenceGenerator(clip, str(tmp_path), RetrievalRefConfig(reference_count=5, similarity_threshold=0.3))
    for t, refs in zip(names[10:], batch[10:]):
        assert [r["index"] for r in refs] == [r["index"] for r in fresh.retrieve_references(t)]
    st = gen.get_statistics()
    assert st["total_queries"] == 40 and st["cache_hits"] == 10 and st["database_info"]["total_references"] == n
