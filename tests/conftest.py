import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def example_points():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "example.bin")
    return np.fromfile(path, dtype=np.float32).reshape(-1, 4)


# a stand-in ground model for the example frame; injected identically everywhere
EXAMPLE_GROUND = [0.0057, -0.0445, -0.9990, -1.7938]
