import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def built_library():
    """Builds libtnsb.so in-tree if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from treensearch_b200 import build
    return build.build_library()


def csr_equal(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
