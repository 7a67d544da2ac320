import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def lib():
    """librfdnet_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from rfdnet_b200 import _lib
    return _lib.load()
