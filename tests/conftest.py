import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs the reference checkout at /root/reference")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    has_cuda = _has_cuda()
    from oracle import ref_import
    has_ref = ref_import.available()
    for item in items:
        if "gpu" in item.keywords and not has_cuda:
            item.add_marker(pytest.mark.skip(reason="no CUDA device in this container"))
        if "needs_reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present (GPU box)"))


@pytest.fixture(scope="session")
def lib():
    """Build (if stale) and load the C-ABI library; GPU tests call through it."""
    from multimodal_uav_det_b200 import build, _lib
    build.build_library()
    return _lib.load()
