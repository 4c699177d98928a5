import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _cuda_library_is_current():
    """On a GPU box make sure the in-tree CUDA library matches the sources (a no-op when it travelled with the snapshot).
    CPU-only runs never build here: tests/test_abi.py asserts the library `__graft_entry__.build()` produced."""
    try:
        import torch

        if torch.cuda.is_available():
            from l4p_b200 import build

            build.build()
    except ImportError:
        pass
    yield


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
