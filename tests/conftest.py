import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both suites (built in-tree by nvcc; no GPU needed to build)."""
    from neighbour_feature_pooling_b200 import build
    build.build(force=False)
    yield


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback to test instead")
    return torch.device("cuda:0")
