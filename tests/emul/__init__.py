"""TEST INFRASTRUCTURE ONLY: ctypes wrapper for the g++ build of builder + per-ray traversal."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(os.path.join(HERE, "libspb_emul.so"))
        L.emul_create.restype = C.c_void_p
        L.emul_create.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int32]
        L.emul_error.restype = C.c_char_p; L.emul_error.argtypes = [C.c_void_p]
        L.emul_destroy.argtypes = [C.c_void_p]
        L.emul_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.emul_sah.restype = C.c_double; L.emul_sah.argtypes = [C.c_void_p]
        L.emul_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
        _lib = L
    return _lib


class Emul:
    def __init__(self, tris, max_leaf=3, bins=32, import_nodes=None):
        tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
        imp = None if import_nodes is None else np.ascontiguousarray(import_nodes)
        self.h = lib().emul_create(tris.ctypes.data, tris.shape[0], max_leaf, bins,
                                   None if imp is None else imp.ctypes.data,
                                   0 if imp is None else len(imp), 0)
        err = lib().emul_error(self.h).decode()
        if err:
            raise RuntimeError(err)
        st = np.zeros(4, np.int64)
        lib().emul_stats(self.h, st.ctypes.data)
        self.n_wide, self.tri_format, self.max_depth, self.n_binary = [int(x) for x in st]
        self.sah = lib().emul_sah(self.h)

    def trace(self, rays, any_hit=False, threads=os.cpu_count() or 1, mode=0):
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        prim = np.empty(n, np.int32); t = np.empty(n, np.float64)
        u = np.empty(n, np.float32); v = np.empty(n, np.float32)
        ctr = np.zeros(3, np.uint64)
        lib().emul_trace(self.h, rays.ctypes.data, int(rays.dtype == np.float64), n, int(any_hit),
                         prim.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data, ctr.ctypes.data, threads, mode)
        self.node_visits = int(ctr[0]) / max(n, 1)
        self.tri_tests = int(ctr[1]) / max(n, 1)
        self.exact_tests = int(ctr[2]) / max(n, 1)
        return prim, t, u, v

    def __del__(self):
        try:
            lib().emul_destroy(self.h)
        except Exception:
            pass
