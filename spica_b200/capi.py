"""ctypes binding of the C ABI (include/spica_b200.h) - what tests/ and bench.py call.

This is plumbing: no algorithm lives here, and there is no fallback.  If libspica_b200.so is
missing, or there is no CUDA device, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libspica_b200.so")

RAY_F32 = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
RAY_F64 = np.dtype([("o", "<f8", 3), ("d", "<f8", 3), ("tmin", "<f8"), ("tmax", "<f8")])
HIT = np.dtype([("t", "<f4"), ("prim", "<i4"), ("u", "<f4"), ("v", "<f4")])
HIT_F64 = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("prim", "<i4"), ("pad", "<i4")])
IMPORT_NODE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("left", "<i4"), ("right", "<i4"),
                        ("prim", "<i4"), ("axis", "<i4")])
assert RAY_F32.itemsize == 32 and RAY_F64.itemsize == 64 and HIT.itemsize == 16
assert HIT_F64.itemsize == 32 and IMPORT_NODE.itemsize == 64


class BuildOpts(C.Structure):
    _fields_ = [("builder", C.c_int32), ("max_leaf_tris", C.c_int32), ("sah_bins", C.c_int32),
                ("reserved_", C.c_int32)]


class BvhStats(C.Structure):
    _fields_ = [("n_tris", C.c_int64), ("n_wide_nodes", C.c_int64), ("n_binary_nodes", C.c_int64),
                ("node_bytes", C.c_int64), ("tri_bytes", C.c_int64), ("sah_cost", C.c_double),
                ("build_seconds", C.c_double), ("tri_format", C.c_int32), ("max_depth", C.c_int32),
                ("world_lo", C.c_double * 3), ("world_hi", C.c_double * 3)]


class Counters(C.Structure):
    _fields_ = [("last_kernel_ms", C.c_double), ("kernel_launches", C.c_int64), ("rays", C.c_int64),
                ("node_visits", C.c_int64), ("tri_tests", C.c_int64)]


class SpbError(RuntimeError):
    pass


_lib = None

# every symbol include/spica_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "spb_version", "spb_last_error", "spb_ctx_create", "spb_ctx_destroy",
    "spb_scene_set_triangles", "spb_bvh_build", "spb_bvh_import_binary", "spb_bvh_get_stats",
    "spb_trace_closest", "spb_trace_closest_f64", "spb_trace_any", "spb_trace_any_f64",
    "spb_trace_closest_dev", "spb_trace_any_dev", "spb_get_counters", "spb_set_option",
    "spb_dev_alloc", "spb_dev_free", "spb_dev_upload", "spb_dev_download", "spb_dev_sync",
    "spb_ctx_stream",
]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpbError("libspica_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.spb_version.restype = C.c_int
    L.spb_last_error.restype = C.c_char_p; L.spb_last_error.argtypes = [vp]
    L.spb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.spb_ctx_destroy.argtypes = [vp]; L.spb_ctx_destroy.restype = None
    L.spb_scene_set_triangles.argtypes = [vp, vp, vp, vp, vp, i64]
    L.spb_bvh_build.argtypes = [vp, C.POINTER(BuildOpts)]
    L.spb_bvh_import_binary.argtypes = [vp, vp, i64, i32]
    L.spb_bvh_get_stats.argtypes = [vp, C.POINTER(BvhStats)]
    for name in ("spb_trace_closest", "spb_trace_closest_f64", "spb_trace_any", "spb_trace_any_f64",
                 "spb_trace_closest_dev", "spb_trace_any_dev"):
        getattr(L, name).argtypes = [vp, vp, i64, vp]
    L.spb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.spb_set_option.argtypes = [vp, C.c_char_p, i64]
    L.spb_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.spb_dev_free.argtypes = [vp, vp]
    L.spb_dev_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.spb_dev_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.spb_dev_sync.argtypes = [vp]
    L.spb_ctx_stream.argtypes = [vp]; L.spb_ctx_stream.restype = vp
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):          # torch tensor (pinned host or device memory)
        return a.data_ptr()
    return a.ctypes.data


class Context:
    """One per GPU (spb_ctx)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.spb_ctx_create(device, C.byref(h))
        if rc != 0:
            raise SpbError("spb_ctx_create failed (%d): %s" % (rc, self.L.spb_last_error(None).decode()))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.spb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SpbError("spica_b200 error %d: %s" % (rc, self.L.spb_last_error(self.h).decode()))

    # ---- scene / BVH
    def set_triangles(self, tris, normals=None, material_id=None, light_id=None):
        tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
        n = tris.shape[0]
        nm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(n, 9)
        mi = None if material_id is None else np.ascontiguousarray(material_id, dtype=np.int32)
        li = None if light_id is None else np.ascontiguousarray(light_id, dtype=np.int32)
        self._check(self.L.spb_scene_set_triangles(self.h, _ptr(tris), _ptr(nm), _ptr(mi), _ptr(li), n))

    def build(self, max_leaf_tris=3, sah_bins=32, builder=0):
        o = BuildOpts(builder, max_leaf_tris, sah_bins, 0)
        self._check(self.L.spb_bvh_build(self.h, C.byref(o)))

    def import_binary(self, nodes, root=0):
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 64
        self._check(self.L.spb_bvh_import_binary(self.h, _ptr(nodes), len(nodes), root))

    def stats(self):
        s = BvhStats()
        self._check(self.L.spb_bvh_get_stats(self.h, C.byref(s)))
        return {k: (list(getattr(s, k)) if k.startswith("world") else getattr(s, k)) for k, _ in s._fields_}

    def set_option(self, name, value):
        self._check(self.L.spb_set_option(self.h, name.encode(), int(value)))

    def counters(self):
        c = Counters()
        self._check(self.L.spb_get_counters(self.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}

    # ---- host-buffer entry points (numpy arrays or pinned torch tensors)
    @staticmethod
    def _rays(rays):
        if hasattr(rays, "data_ptr"):
            n = rays.shape[0]
            f64 = rays.element_size() == 8
            return rays, n, f64
        rays = np.ascontiguousarray(rays)
        if rays.dtype in (RAY_F32, RAY_F64):
            return rays, rays.shape[0], rays.dtype == RAY_F64
        assert rays.ndim == 2 and rays.shape[1] == 8 and rays.dtype in (np.float32, np.float64)
        return rays, rays.shape[0], rays.dtype == np.float64

    def trace_closest(self, rays, out=None):
        rays, n, f64 = self._rays(rays)
        if out is None:
            out = np.empty(n, dtype=HIT_F64 if f64 else HIT)
        fn = self.L.spb_trace_closest_f64 if f64 else self.L.spb_trace_closest
        self._check(fn(self.h, _ptr(rays), n, _ptr(out)))
        return out

    def trace_any(self, rays, out=None):
        rays, n, f64 = self._rays(rays)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        fn = self.L.spb_trace_any_f64 if f64 else self.L.spb_trace_any
        self._check(fn(self.h, _ptr(rays), n, _ptr(out)))
        return out

    # ---- device-resident entry points (raw device pointers or torch cuda tensors)
    def trace_closest_dev(self, d_rays, n, d_hits):
        self._check(self.L.spb_trace_closest_dev(self.h, _ptr(d_rays), n, _ptr(d_hits)))

    def trace_any_dev(self, d_rays, n, d_occ):
        self._check(self.L.spb_trace_any_dev(self.h, _ptr(d_rays), n, _ptr(d_occ)))

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.spb_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        self._check(self.L.spb_dev_free(self.h, p))

    def dev_upload(self, d_dst, host):
        host = np.ascontiguousarray(host)
        self._check(self.L.spb_dev_upload(self.h, d_dst, _ptr(host), host.nbytes))

    def dev_download(self, host, d_src):
        self._check(self.L.spb_dev_download(self.h, _ptr(host), d_src, host.nbytes))

    def sync(self):
        self._check(self.L.spb_dev_sync(self.h))
