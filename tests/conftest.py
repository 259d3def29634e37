import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test, enabled with VT_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("VT_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="set VT_SLOW=1 to run")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def vt():
    import vtree_b200
    return vtree_b200


@pytest.fixture(scope="session")
def cuda(vt):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    vt._native.require_device()          # raises if the native library is missing: no silent fallback
    return torch.device("cuda:0")
