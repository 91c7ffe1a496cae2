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


def transformed_torus_case(offset, n_incoherent=40000, n_aimed=20000):
    """A torus under a double-precision toWorld transform (vertices are not float32 numbers; `offset` moves it away from the
    origin so that the float32 rounding step is large against the triangles) + rays: incoherent ones and ones aimed at
    mesh vertices and edge midpoints.  Returns (tris float64 [n, 9], rays float32 [m, 8], lo, hi)."""
    import numpy as np
    from spica_b200 import scenes
    v, f = scenes.torus_mesh(160, 80)
    c, s_ = np.cos(0.7), np.sin(0.7)
    rot = np.array([[c, -s_, 0], [s_, c * np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    vd = v.astype(np.float64) @ rot.T * 1.37 + np.array([offset, -0.5 * offset, 0.1])
    tris = np.ascontiguousarray(vd[f].reshape(len(f), 9))
    lo, hi = vd.min(0), vd.max(0)
    rng = np.random.default_rng(12)
    pick = rng.integers(0, len(f), n_aimed)
    h = n_aimed // 2
    tgt = np.concatenate([vd[f[pick[:h], 0]], 0.5 * (vd[f[pick[h:], 0]] + vd[f[pick[h:], 1]])])
    aimed = np.zeros((n_aimed, 8), np.float32)
    aimed[:, :3] = lo + (hi - lo) * rng.uniform(-0.2, 1.2, (n_aimed, 3))
    aimed[:, 3:6] = tgt - aimed[:, :3].astype(np.float64)
    aimed[:, 7] = 1e32
    rays = np.concatenate([scenes.incoherent_rays(n_incoherent, lo, hi, seed=11), aimed])
    return tris, rays, lo, hi


def random_soup_cases(seed, n_random=3000, n_aimed=3000):
    """Random triangle soups (needles, slivers, overlapping and touching triangles) at a random scale and offset: yields
    (tris float64 [n, 9], rays float32 [m, 8]) twice -- float32-exact vertices, then arbitrary doubles -- with random rays and
    rays aimed at vertices, edges and interior points of the triangles."""
    import numpy as np
    from spica_b200 import scenes
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 400))
    scale = 10.0 ** rng.uniform(-3, 3)
    offset = rng.uniform(-1, 1, 3) * scale * 10.0 ** rng.uniform(-1, 2.5)
    base = rng.uniform(-1, 1, (n, 1, 3))
    ext = 10.0 ** rng.uniform(-3, 0, (n, 1, 1))
    shape = rng.normal(size=(n, 3, 3)) * ext
    k = n // 5                                                                     # slivers
    shape[:k, 2] = shape[:k, 0] + (shape[:k, 1] - shape[:k, 0]) * rng.uniform(0, 1, (k, 1)) + rng.normal(size=(k, 3)) * 1e-6
    v64 = (base + shape) * scale + offset
    for tris in (v64.astype(np.float32).astype(np.float64).reshape(n, 9), v64.reshape(n, 9)):
        lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
        m = n_aimed
        tv = tris.reshape(n, 3, 3)[rng.integers(0, n, m)]
        w = rng.dirichlet([0.3, 0.3, 0.3], m)                                    # near vertices and edges more often than not
        w[: m // 6] = np.eye(3)[rng.integers(0, 3, m // 6)]                        # exactly at a vertex
        tgt = (tv * w[:, :, None]).sum(1)
        aimed = np.zeros((m, 8), np.float32)
        aimed[:, :3] = lo + (hi - lo) * rng.uniform(-0.5, 1.5, (m, 3))
        aimed[:, 3:6] = tgt - aimed[:, :3].astype(np.float64)
        aimed[:, 7] = 1e32
        yield tris, np.concatenate([scenes.incoherent_rays(n_random, lo, hi, seed=seed), aimed])
