"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU restatement (oracle/spica_oracle.cc) and
subprocess access to the compiled unmodified reference (oracle/_ref/*).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (spica_b200/) never does.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libspica_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
RAYCAST_REF = os.path.join(REF_DIR, "raycast_ref")
RENDER_REF = os.path.join(REF_DIR, "render_ref")
SPICA_REF = os.path.join(REF_DIR, "bin", "spica")

NODE_DTYPE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("left", "<i4"), ("right", "<i4"),
                       ("prim", "<i4"), ("axis", "<i4")])
assert NODE_DTYPE.itemsize == 64


def build_port():
    """Compile the restatement (gcc only; no reference needed)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def build_ref():
    """Compile the unmodified reference + harnesses when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            build_port()
        L = C.CDLL(PORT_SO)
        dp, vp, i64, i32 = C.POINTER(C.c_double), C.c_void_p, C.c_int64, C.c_int32
        L.so_ray_init.argtypes = [dp, dp, dp, dp]; L.so_ray_init.restype = C.c_int
        L.so_triangle_intersect.argtypes = [dp, dp, dp, C.c_double, dp, dp, dp]
        L.so_triangle_intersect.restype = C.c_int
        L.so_bounds_intersect.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp]
        L.so_bounds_intersect.restype = C.c_int
        L.so_bvh_build.argtypes = [vp, i64, vp]; L.so_bvh_build.restype = i64
        L.so_trace_closest.argtypes = [vp, i32, vp, vp, C.c_int, i64, vp, vp, vp, vp, C.c_int]
        L.so_trace_any.argtypes = [vp, i32, vp, vp, C.c_int, i64, vp, C.c_int]
        L.so_trace_bruteforce.argtypes = [vp, i64, vp, C.c_int, i64, vp, vp, C.c_int]
        L.so_count_ordered_visits.argtypes = [vp, i32, vp, vp, C.c_int, i64, vp, vp, C.c_int]
        _lib = L
    return _lib


def _d3(a):
    return (C.c_double * len(a))(*[float(x) for x in a])


def ray_init(o, d):
    dir_ = (C.c_double * 3)(); inv = (C.c_double * 3)()
    ok = lib().so_ray_init(_d3(o), _d3(d), dir_, inv)
    return bool(ok), list(dir_), list(inv)


def triangle_intersect(tri9, org, dir_, max_dist=1.0e32):
    t, u, v = C.c_double(), C.c_double(), C.c_double()
    ok = lib().so_triangle_intersect(_d3(tri9), _d3(org), _d3(dir_), max_dist,
                                     C.byref(t), C.byref(u), C.byref(v))
    return bool(ok), t.value, u.value, v.value


def bounds_intersect(lo, hi, org, invdir, max_dist=1.0e32):
    tn, tf = C.c_double(), C.c_double()
    ok = lib().so_bounds_intersect(_d3(lo), _d3(hi), _d3(org), _d3(invdir), max_dist,
                                   C.byref(tn), C.byref(tf))
    return bool(ok), tn.value, tf.value


def _rays(rays):
    rays = np.ascontiguousarray(rays)
    assert rays.ndim == 2 and rays.shape[1] == 8 and rays.dtype in (np.float32, np.float64)
    return rays, int(rays.dtype == np.float64)


def bvh_build(tris):
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
    n = tris.shape[0]
    nodes = np.zeros(max(2 * n - 1, 0), dtype=NODE_DTYPE)
    cnt = lib().so_bvh_build(tris.ctypes.data, n, nodes.ctypes.data)
    return nodes[:cnt]


def trace_closest(nodes, tris, rays, threads=os.cpu_count() or 1, root=0):
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
    rays, f64 = _rays(rays)
    n = rays.shape[0]
    prim = np.empty(n, np.int32); t = np.empty(n, np.float64)
    u = np.empty(n, np.float64); v = np.empty(n, np.float64)
    nodes = np.ascontiguousarray(nodes)
    lib().so_trace_closest(nodes.ctypes.data, root if len(nodes) else -1, tris.ctypes.data,
                           rays.ctypes.data, f64, n, prim.ctypes.data, t.ctypes.data,
                           u.ctypes.data, v.ctypes.data, threads)
    return prim, t, u, v


def trace_any(nodes, tris, rays, threads=os.cpu_count() or 1, root=0):
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
    rays, f64 = _rays(rays)
    n = rays.shape[0]
    occ = np.empty(n, np.uint8)
    nodes = np.ascontiguousarray(nodes)
    lib().so_trace_any(nodes.ctypes.data, root if len(nodes) else -1, tris.ctypes.data,
                       rays.ctypes.data, f64, n, occ.ctypes.data, threads)
    return occ


def trace_bruteforce(tris, rays, threads=os.cpu_count() or 1):
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
    rays, f64 = _rays(rays)
    n = rays.shape[0]
    prim = np.empty(n, np.int32); t = np.empty(n, np.float64)
    lib().so_trace_bruteforce(tris.ctypes.data, tris.shape[0], rays.ctypes.data, f64, n,
                              prim.ctypes.data, t.ctypes.data, threads)
    return prim, t


def count_ordered_visits(nodes, tris, rays, threads=os.cpu_count() or 1, root=0):
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
    rays, f64 = _rays(rays)
    n = rays.shape[0]
    a, b = C.c_int64(), C.c_int64()
    nodes = np.ascontiguousarray(nodes)
    lib().so_count_ordered_visits(nodes.ctypes.data, root, tris.ctypes.data, rays.ctypes.data,
                                  f64, n, C.byref(a), C.byref(b), threads)
    return a.value / max(n, 1), b.value / max(n, 1)


# ---------------------------------------------------------------------------------------------
# the compiled, unmodified reference
# ---------------------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(RAYCAST_REF)


def ref_raycast(rays, mode="closest", ply=None, tris=None, threads=None, simd=0, stride=1,
                dump_bvh=False, repeat=1, warmup=0):
    """Run oracle/_ref/raycast_ref. Returns (info, prim, t) | (info, occluded) [+ nodes]."""
    assert have_ref(), "oracle/_ref/raycast_ref missing: run `make -C oracle ref` where /root/reference exists"
    with tempfile.TemporaryDirectory() as td:
        cmd = [RAYCAST_REF, "--mode", mode, "--simd", str(simd), "--stride", str(stride),
               "--repeat", str(repeat), "--warmup", str(warmup)]
        if threads:
            cmd += ["--threads", str(threads)]
        if ply is not None:
            cmd += ["--ply", ply]
        else:
            tp = os.path.join(td, "tris.f64")
            np.ascontiguousarray(tris, dtype=np.float64).tofile(tp)
            cmd += ["--tris", tp]
        n = 0
        if rays is not None:
            rays = np.ascontiguousarray(rays)
            rp = os.path.join(td, "rays.bin")
            rays.tofile(rp)
            cmd += ["--rays", rp, "--ray-format", "f64" if rays.dtype == np.float64 else "f32",
                    "--out", os.path.join(td, "out.bin")]
            n = len(range(0, rays.shape[0], stride))
        if dump_bvh:
            cmd += ["--dump-bvh", os.path.join(td, "bvh.bin")]
        out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
        info = json.loads(out.strip().splitlines()[-1])
        res = [info]
        if rays is not None:
            raw = np.fromfile(os.path.join(td, "out.bin"), dtype=np.uint8)
            if mode == "closest":
                res.append(raw[:4 * n].view(np.int32).copy())
                res.append(raw[4 * n:12 * n].view(np.float64).copy())
            else:
                res.append(raw[:n].copy())
        if dump_bvh:
            raw = np.fromfile(os.path.join(td, "bvh.bin"), dtype=np.uint8)
            hdr = raw[:8].view(np.int32)
            nodes = raw[8:8 + 64 * int(hdr[0])].view(NODE_DTYPE).copy()
            assert int(hdr[1]) in (0, -1)
            res.append(nodes)
        return tuple(res)
