import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def has_cuda_device():
    """True when a context can be created.  (Through the C ABI, not torch: importing torch AFTER this repo's CUDA libraries
    were loaded into the process fails, and tests may run in any subset and order.)"""
    from spica_b200 import capi
    try:
        capi.Context(0).close()
        return True
    except capi.SpbError:
        return False


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_torus():
    g = load_golden("raycast_torus_1600.npz")
    g["tris"] = g["verts"].astype(np.float64)[g["faces"]].reshape(-1, 9)
    return g


@pytest.fixture(scope="session")
def golden_cube():
    g = load_golden("raycast_cube_12.npz")
    g["tris"] = g["verts"].astype(np.float64)[g["faces"]].reshape(-1, 9)
    return g


@pytest.fixture(scope="session")
def golden_f64verts():
    return load_golden("raycast_torus_f64verts.npz")


@pytest.fixture(scope="session")
def gpu_ctx():
    """A live context on cuda:0 through the C ABI. No CPU fallback: fails if the library is absent."""
    from spica_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
