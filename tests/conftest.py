import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built library (it is git-ignored): build it once, in-tree (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "jaxrk_b200", "lib", "libjaxrk_b200.so")
    if not os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        from jaxrk_b200 import build as _b
        _b.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_numpy_shim.npz"))


@pytest.fixture(scope="session")
def c_oracle():
    """ctypes handle on oracle/liboracle.so (built on demand; test infrastructure only)."""
    from oracle import c_oracle as co
    return co.load()
