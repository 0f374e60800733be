import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def tvc_ctx():
    import multimodal_detection_consistency_b200 as tvc
    return tvc.Context.get(0)
