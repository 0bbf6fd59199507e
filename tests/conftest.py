import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without CUDA: skip instead of failing with a confusing error
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build (if stale) and return the path of the C-ABI library; CPU-only, no compute."""
    from canonicalvoting_b200 import build
    return build.build()


@pytest.fixture(autouse=True)
def _exact_fp32_unless_a_test_says_otherwise():
    """The parity tests compare the module path with the oracle at fp32 accuracy: pin the exact-fp32 convolution for every test
    (the library default is 'auto': tensor cores wherever autograd is off).  Tests of the tensor-core path select it themselves."""
    from canonicalvoting_b200.sparse import functional
    functional.set_forward_mode("fp32")
    yield
    functional.set_forward_mode("fp32")
