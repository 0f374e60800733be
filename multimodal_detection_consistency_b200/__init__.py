"""B200-native TVC scoring + retrieval hot path (libtvc.so + the reference-shaped host API).

Reference-shaped classes (same names / configs / result shapes as the reference's modules):
  retrieval.MultiModalRetriever, RetrievalConfig, FaissIndexManager, RetrievalIndex, ConsistencyCalculator
  ref_bank.ReferenceBank, ReferenceBankConfig, ReferenceItem
  detector.AdversarialDetector, DetectorConfig
  defenses.ConsistencyChecker, MultiModalDefenseDetector, DetectionConfig, RetrievalReferenceIndex,
           RetrievalReferenceGenerator, RetrievalRefConfig
  hubness.compute_hubness, k_occurrence, hubness_scores
  metrics.RetrievalEvaluator, RetrievalMetrics, SimilarityCalculator;  batching.MicroBatcher
  faiss_compat  (install as sys.modules["faiss"] to route the reference's own files here)
Batched engine: pipeline.TVCScorer;  raw kernels: Context, Gallery.
"""
from ._native import (Context, DetectorParams, Gallery, TvcError, default_params, load_library,  # noqa: F401
                      EXPORTED_SYMBOLS, SCORE_NAMES, SCORE_INDEX, NSCORES)

__version__ = "0.1.0"

_LAZY = {
    "MultiModalRetriever": "retrieval", "RetrievalConfig": "retrieval", "FaissIndexManager": "retrieval",
    "RetrievalIndex": "retrieval", "ConsistencyCalculator": "retrieval", "RetrievalResult": "retrieval",
    "IndexConfig": "retrieval", "create_retriever": "retrieval",
    "ReferenceBank": "ref_bank", "ReferenceBankConfig": "ref_bank", "ReferenceItem": "ref_bank",
    "create_reference_bank": "ref_bank",
    "AdversarialDetector": "detector", "DetectorConfig": "detector", "create_adversarial_detector": "detector",
    "ConsistencyChecker": "defenses", "MultiModalDefenseDetector": "defenses", "DetectionConfig": "defenses",
    "RetrievalReferenceIndex": "defenses", "RetrievalReferenceGenerator": "defenses", "RetrievalRefConfig": "defenses",
    "compute_hubness": "hubness", "k_occurrence": "hubness", "hubness_scores": "hubness",
    "compute_hubness_loss": "hubness",
    "TVCScorer": "pipeline",
    "RetrievalEvaluator": "metrics", "RetrievalMetrics": "metrics", "SimilarityCalculator": "metrics",
    "SimilarityMetrics": "metrics",
    "MicroBatcher": "batching",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
