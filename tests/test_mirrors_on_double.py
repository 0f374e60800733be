"""Host logic of the Python mirrors without a GPU: the golden-fixture API tests of tests/test_gpu_api.py
(detector, ConsistencyChecker, defense detector, ReferenceBank, retriever) re-run with the native layer
replaced by the oracle-backed test double (tests/fake_native.py).  What the GPU run adds on top is the
kernels; what this run pins, on any machine, is everything above the C ABI against the reference's
recorded outputs (tests/golden/*.npz)."""
import importlib.util
from pathlib import Path

import pytest

import fake_native

HERE = Path(__file__).resolve().parent
NAMES = ["test_adversarial_detector_matches_reference_outputs", "test_consistency_checker_matches_reference_outputs",
         "test_defense_detector_batched_matches_oracle", "test_reference_bank_matches_reference_outputs",
         "test_reference_bank_insert_dedup_eviction_persistence", "test_retriever_drop_in",
         "test_reference_bank_journal_never_drifts_from_memory", "test_reference_bank_kmeans_assignment_matches_sklearn"]


@pytest.fixture(scope="module")
def api_tests():
    spec = importlib.util.spec_from_file_location("gpu_api_tests_on_double", HERE / "test_gpu_api.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", NAMES)
def test_api_test_passes_on_the_double(api_tests, name, tmp_path):
    import inspect
    fn = getattr(api_tests, name)
    with fake_native.installed() as ctx:
        kwargs = {p: (ctx if p == "tvc_ctx" else tmp_path) for p in inspect.signature(fn).parameters}
        assert set(kwargs) <= {"tvc_ctx", "tmp_path"}
        fn(**kwargs)
