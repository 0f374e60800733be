"""Drop-in check under the reference's OWN orchestrator (build container only: needs /root/reference).

`src/pipeline.py` (MultiModalDetectionPipeline: process_single, process_batch on its 4 worker threads,
clear_cache, get_stats) is imported unmodified and run twice on the same seeded table encoders:

  (A) with the reference's MultiModalRetriever / AdversarialDetector.  faiss is not installable here, so
      `faiss.IndexFlatIP` is a 15-line NumPy stand-in (exact inner product, descending) - the published
      contract the reference relies on (src/retrieval.py:495-518, 652-656);
  (C) with the reference's own classes and only the `faiss` module replaced by faiss_compat (INTEGRATION.md
      level 1: no source change at all);
  (B) with this repo's mirrors swapped in exactly as INTEGRATION.md level 2 says - the two import lines
      of src/pipeline.py:19,21 - and NOTHING else changed.  This container has no GPU, so the native
      layer under the mirrors is the oracle-backed test double tests/fake_native.py; what is exercised
      is everything above the C ABI: construction from the pipeline's config objects, lazy encoder
      acquisition, return shapes, result-dict keys, error convention, caches, stats, thread safety and
      micro-batching of the concurrent single-sample calls.

Every PipelineResult field the path produces must agree: retrieved paths (order), retrieval scores,
is_adversarial, detection score, the nested detection_details, pipeline steps, no errors.

    python tests/golden/pipeline_dropin.py [seed ...]
"""
from __future__ import annotations

import importlib
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
from PIL import Image

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))

import make_golden as MG  # noqa: E402
import fake_native  # noqa: E402

D, N_GALLERY, N_SAMPLES, V, G = 64, 80, 24, 5, 3


class NumpyFlatIP:
    """faiss.IndexFlatIP stand-in for arm (A): exact inner product, scores descending."""

    def __init__(self, d, *a):
        self.d, self.x, self.is_trained = d, np.zeros((0, d), np.float32), True

    ntotal = property(lambda self: len(self.x))

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, np.float32)])

    def search(self, q, k):
        s = np.asarray(q, np.float32) @ self.x.T
        idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
        return np.take_along_axis(s, idx, 1), idx.astype(np.int64)


def image_of(i: int) -> Image.Image:
    """A 2x2 image whose first pixel encodes the id of its embedding."""
    return Image.new("RGB", (2, 2), (i % 256, (i // 256) % 256, i // 65536))


class TableClip:
    """Encoder stand-in at the boundary north_star draws: embeddings are looked up, not computed."""

    def __init__(self, text_table, image_table):
        self.t, self.i, self.text_calls, self.image_calls = text_table, image_table, 0, 0

    @staticmethod
    def _id(im):
        if isinstance(im, torch.Tensor):
            return int(im.reshape(-1)[0])
        r, g, b = im.convert("RGB").getpixel((0, 0))
        return r + 256 * g + 65536 * b

    def encode_text(self, texts, normalize=True):
        self.text_calls += 1
        texts = [texts] if isinstance(texts, str) else list(texts)
        return torch.stack([torch.from_numpy(self.t[s]) for s in texts])

    def encode_image(self, images, normalize=True):
        self.image_calls += 1
        images = list(images) if isinstance(images, (list, tuple)) else [images]
        return torch.stack([torch.from_numpy(self.i[self._id(im)]) for im in images])

    def get_text_image_similarity(self, text, image):
        a, b = self.encode_text([text]), self.encode_image([image])
        return torch.nn.functional.cosine_similarity(a, b)[0]


def build_world(seed: int, root: Path):
    rng = np.random.default_rng(seed)
    sd = 1.0 / np.sqrt(D)
    gal = MG.unit(rng, N_GALLERY, D)
    paths = []
    image_table, text_table = {}, {}
    for j in range(N_GALLERY):
        p = root / f"gallery_{j:03d}.png"
        image_of(j).save(p)
        paths.append(str(p))
        image_table[j] = gal[j]
    samples = []
    for i in range(N_SAMPLES):
        pick = int(rng.integers(N_GALLERY))
        text = f"a photo matching gallery item {pick} (query {i})"
        t = gal[pick] + 0.5 * sd * rng.standard_normal(D).astype(np.float32)
        text_table[text] = (t / np.linalg.norm(t)).astype(np.float32)
        for v in range(V):
            tv = text_table[text] + 0.25 * sd * rng.standard_normal(D).astype(np.float32)
            text_table[f"{text} ~v{v}"] = (tv / np.linalg.norm(tv)).astype(np.float32)
        img_id = 1000 + i
        if rng.uniform() < 0.4:                                    # attacked: image unrelated to its text
            im = MG.unit(rng, 1, D)[0] + 0.3 * gal[int(rng.integers(N_GALLERY))]
        else:
            im = gal[pick] + 0.5 * sd * rng.standard_normal(D).astype(np.float32)
        image_table[img_id] = (im / np.linalg.norm(im)).astype(np.float32)
        for g in range(G):
            ref = 0.6 * image_table[img_id] + 2.4 * sd * rng.standard_normal(D).astype(np.float32)
            image_table[5000 + G * i + g] = (ref / np.linalg.norm(ref)).astype(np.float32)
        samples.append((image_of(img_id), text, i))
    return gal, paths, text_table, image_table, samples


def run_arm(P, world, swap):
    gal, paths, text_table, image_table, samples = world
    clip = TableClip(text_table, image_table)
    index_of = {s[1]: s[2] for s in samples}
    augmenter = types.SimpleNamespace(generate_variants=lambda text: [f"{text} ~v{v}" for v in range(V)])
    sd_gen = types.SimpleNamespace(generate_reference_images=lambda text, num_images=G: {
        "images": [image_of(5000 + G * index_of[text] + g) for g in range(num_images)], "generation_time": 0.0})
    cfg = P.PipelineConfig(enable_text_augment=False, enable_sd_reference=False, enable_retrieval=True,
                           enable_detection=True, enable_parallel=True, max_workers=4)
    pipe = P.MultiModalDetectionPipeline(cfg)                       # builds retriever + detector from the configs
    # encoders: the attributes src/retrieval.py:338 and src/detector.py:227-229 hold
    pipe.retriever.clip_model = clip
    pipe.detector.clip_model, pipe.detector.text_augmenter, pipe.detector.sd_generator = clip, augmenter, sd_gen
    feats = pipe.retriever.build_image_index(paths)                 # PIL load -> encoder -> index (src/retrieval.py:371)
    assert np.allclose(np.asarray(feats), gal, atol=0) and pipe.retriever.image_index is not None
    first = pipe.process_single(samples[0][0], samples[0][1])       # sequential caller
    batch = pipe.process_batch([s[0] for s in samples], [s[1] for s in samples])   # 4 worker threads
    again = pipe.process_single(samples[0][0], samples[0][1])       # served from both caches
    stats = pipe.get_stats()
    pipe.clear_cache()
    return dict(first=first, batch=sorted(batch, key=lambda r: r.original_text), again=again, stats=stats,
                clip=clip, pipe=pipe)


def same(a, b, where, tol=1e-5):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), (where, sorted(a), sorted(b))
        for k in a:
            if k not in ("detection_time", "generation_time"):
                same(a[k], b[k], f"{where}.{k}", tol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (where, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{where}[{i}]", tol)
    elif isinstance(a, (bool, np.bool_, str, type(None))):
        assert a == b, (where, a, b)
    else:
        assert abs(float(a) - float(b)) <= tol, (where, a, b)


def compare(ra, rb, where):
    assert ra.errors == [] and rb.errors == [], (where, ra.errors, rb.errors)
    assert ra.original_text == rb.original_text and ra.pipeline_steps == rb.pipeline_steps
    assert [np.array(x).tolist() for x in ra.retrieved_images] == [np.array(x).tolist() for x in rb.retrieved_images], where
    assert len(ra.retrieved_images) == 5 and ra.retrieved_texts == rb.retrieved_texts
    same(list(ra.retrieval_scores), list(rb.retrieval_scores), where + ".retrieval_scores")
    assert bool(ra.is_adversarial) == bool(rb.is_adversarial), (where, ra.detection_score, rb.detection_score)
    same(ra.detection_score, rb.detection_score, where + ".detection_score")
    same(ra.detection_details, rb.detection_details, where + ".detection_details")
    same({k: v for k, v in ra.to_dict().items() if not k.endswith("_time")},
         {k: v for k, v in rb.to_dict().items() if not k.endswith("_time")}, where + ".to_dict")


def main():
    seeds = [int(s) for s in sys.argv[1:]] or [7, 8]
    MG.import_reference()
    sys.modules["faiss"].IndexFlatIP = NumpyFlatIP
    P = importlib.import_module("src.pipeline")
    # the reference's plotting helper cannot be constructed as shipped (src/utils/visualization.py:1024 names an
    # undefined TSNEVisualizer); plotting is outside the path, so both arms get an inert one
    P.ExperimentVisualizer = lambda *a, **k: types.SimpleNamespace()
    ref_names = {n: getattr(P, n) for n in ("MultiModalRetriever", "RetrievalConfig", "AdversarialDetector",
                                            "DetectorConfig")}
    from multimodal_detection_consistency_b200 import detector as our_det, retrieval as our_ret
    for seed in seeds:
        with tempfile.TemporaryDirectory() as td:
            world = build_world(seed, Path(td))
            for n, v in ref_names.items():
                setattr(P, n, v)
            a = run_arm(P, world, swap=False)
            assert type(a["pipe"].retriever).__module__ == "src.retrieval"
            # ---- the swap INTEGRATION.md describes: the two import lines of src/pipeline.py ----
            P.MultiModalRetriever, P.RetrievalConfig = our_ret.MultiModalRetriever, our_ret.RetrievalConfig
            P.AdversarialDetector, P.DetectorConfig = our_det.AdversarialDetector, our_det.DetectorConfig
            with fake_native.installed() as ctx:
                b = run_arm(P, world, swap=True)
            # ---- arm (C), INTEGRATION.md level 1: the reference's OWN classes, only `faiss` swapped for
            # faiss_compat (sys.modules["faiss"] = faiss_compat before `import src.retrieval`) ----
            for n, v in ref_names.items():
                setattr(P, n, v)
            from multimodal_detection_consistency_b200 import faiss_compat
            fa = sys.modules["faiss"]
            for name in ("IndexFlatIP", "IndexIVFFlat", "IndexHNSWFlat", "StandardGpuResources", "index_cpu_to_gpu",
                         "get_num_gpus", "write_index", "read_index"):
                setattr(fa, name, getattr(faiss_compat, name))
            with fake_native.installed():
                c = run_arm(P, world, swap=False)
                assert type(c["pipe"].retriever).__module__ == "src.retrieval"
                assert isinstance(c["pipe"].retriever.image_index, faiss_compat.IndexFlatIP)
            fa.IndexFlatIP = NumpyFlatIP
            compare(a["first"], c["first"], "faiss_compat first")
            for ra, rc in zip(a["batch"], c["batch"]):
                compare(ra, rc, f"faiss_compat batch[{ra.original_text}]")
            assert type(b["pipe"].retriever).__module__.startswith("multimodal_detection_consistency_b200")
            assert type(b["pipe"].detector).__module__.startswith("multimodal_detection_consistency_b200")
            compare(a["first"], b["first"], "first")
            compare(a["again"], b["again"], "again")
            assert len(a["batch"]) == len(b["batch"]) == N_SAMPLES
            for ra, rb in zip(a["batch"], b["batch"]):
                compare(ra, rb, f"batch[{ra.original_text}]")
            flagged = sum(bool(r.is_adversarial) for r in b["batch"])
            assert 0 < flagged < N_SAMPLES, flagged
            # stats keep the reference's keys
            sa, sb = a["stats"], b["stats"]
            assert sorted(sa) == sorted(sb), (sorted(sa), sorted(sb))
            assert sa["pipeline_stats"]["total_processed"] == sb["pipeline_stats"]["total_processed"] == N_SAMPLES + 2
            assert sb["pipeline_stats"]["failed_processed"] == 0
            for comp in ("retriever_stats", "detector_stats"):
                if comp in sa:
                    assert sorted(sa[comp]) == sorted(sb[comp]), (comp, sorted(sa[comp]), sorted(sb[comp]))
            # the concurrent single-sample calls were coalesced: fewer scoring launches and encoder calls than samples
            print(f"seed {seed}: {N_SAMPLES + 2} samples identical; {flagged} flagged adversarial; arm B scoring launches "
                  f"{ctx.launches} (arm A: one Python scoring pass per sample), text-encoder calls A/B "
                  f"{a['clip'].text_calls}/{b['clip'].text_calls}, image-encoder calls A/B "
                  f"{a['clip'].image_calls}/{b['clip'].image_calls}")
            assert ctx.launches <= N_SAMPLES + 1 and b["clip"].text_calls < a["clip"].text_calls
    print("pipeline drop-in ok")


if __name__ == "__main__":
    main()
