import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def g():
    """The product binding; building is the driver's job (__graft_entry__.build), but make tests self-contained."""
    import gemmul8_b200
    if not os.path.exists(gemmul8_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return gemmul8_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.cpu()
    return o


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz")) if os.path.isdir(GOLDEN_DIR) else []
