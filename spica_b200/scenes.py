"""Synthetic scenes and ray sets named by BASELINE.json.

Everything here is deterministic (counter-based hashes, no global RNG state) so that the CPU
oracle, the compiled reference and the GPU see bit-identical float32 inputs.
"""
import os
import struct

import numpy as np

INFTY = np.float32(1.0e32)  # reference core/common.h:54, as stored in a float32 ray record


def _hash_u32(x):
    """lowbias32-style integer hash on uint32 arrays (wraps mod 2^32)."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def _u01(idx, dim, seed):
    """Counter-based uniform in [0,1): keyed by (ray index, dimension, seed), 24-bit mantissa."""
    idx = np.asarray(idx, dtype=np.uint64)
    k = _hash_u32((idx & 0xFFFFFFFF) ^ np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF))
    k = _hash_u32(k.astype(np.uint64) + np.uint64((dim * 0x85EBCA6B + (seed << 7)) & 0xFFFFFFFF)
                  + (idx >> 32))
    return (k >> 8).astype(np.float64) * (1.0 / 16777216.0)


def torus_mesh(nu=1000, nv=500, R=1.0, r=0.4, bump=0.05, seed=1):
    """Closed displaced torus with exactly 2*nu*nv triangles, float32 vertices.

    Returns (verts float32 [nu*nv,3], faces int32 [2*nu*nv,3])."""
    i = np.arange(nu, dtype=np.int64)[:, None]
    j = np.arange(nv, dtype=np.int64)[None, :]
    h = _hash_u32((i * 73856093 ^ j * 19349663 ^ (seed * 83492791)) & 0xFFFFFFFF)
    h = (h >> 8).astype(np.float64) * (2.0 / 16777216.0) - 1.0
    u = 2.0 * np.pi * i / nu
    v = 2.0 * np.pi * j / nv
    rr = r * (1.0 + bump * h)
    x = (R + rr * np.cos(v)) * np.cos(u)
    y = (R + rr * np.cos(v)) * np.sin(u)
    z = rr * np.sin(v) + 0.0 * u
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    vid = lambda a, b: ((a % nu) * nv + (b % nv))
    a = vid(i, j); b = vid(i + 1, j); c = vid(i + 1, j + 1); d = vid(i, j + 1)
    f0 = np.stack([a, b, c], axis=-1).reshape(-1, 3)
    f1 = np.stack([a, c, d], axis=-1).reshape(-1, 3)
    faces = np.empty((2 * nu * nv, 3), dtype=np.int32)
    faces[0::2] = f0
    faces[1::2] = f1
    return verts, faces


def mesh_triangles(verts, faces, dtype=np.float64):
    """[n,9] world-space triangle soup (p0,p1,p2) - exact widening of the float32 vertices."""
    return verts.astype(dtype)[faces].reshape(-1, 9)


def write_ply(path, verts, faces):
    """binary_little_endian PLY the reference's loader accepts (core/meshio.cc:76-166)."""
    verts = np.ascontiguousarray(verts, dtype="<f4")
    faces = np.ascontiguousarray(faces, dtype="<i4")
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n"
           "property float x\nproperty float y\nproperty float z\n"
           "element face %d\nproperty list uchar int vertex_indices\nend_header\n"
           % (len(verts), len(faces)))
    rec = np.empty(len(faces), dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    rec["n"] = 3
    rec["i"] = faces
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(verts.tobytes())
        f.write(rec.tobytes())


def read_ply(path):
    """Minimal reader for the same PLY subset (float xyz [+ ignored extra float props], tri faces)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    nv = nf = 0
    nprops = 0
    section = None
    for line in data[:end].decode("ascii", "replace").splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "element":
            section = t[1]
            if t[1] == "vertex":
                nv = int(t[2])
            elif t[1] == "face":
                nf = int(t[2])
        elif t[0] == "property" and section == "vertex":
            nprops += 1
    verts = np.frombuffer(data, dtype="<f4", count=nv * nprops, offset=end).reshape(nv, nprops)[:, :3]
    off = end + nv * nprops * 4
    rec = np.frombuffer(data, dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]), count=nf, offset=off)
    assert (rec["n"] == 3).all(), "only triangle faces supported"
    return np.array(verts, dtype=np.float32), np.array(rec["i"], dtype=np.int32)


def primary_rays(width=4096, height=4096, eye=(0.0, 0.0, 4.0), fov_deg=40.0):
    """Pinhole at `eye` looking at the origin (-z), one ray per pixel centre; float32 [n,8]."""
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64),
                         indexing="ij")
    th = np.tan(np.deg2rad(fov_deg) * 0.5)
    aspect = width / height
    px = ((xs + 0.5) / width * 2.0 - 1.0) * th * aspect
    py = (1.0 - (ys + 0.5) / height * 2.0) * th
    rays = np.zeros((width * height, 8), dtype=np.float32)
    rays[:, 0:3] = np.asarray(eye, dtype=np.float32)
    rays[:, 3] = px.reshape(-1)
    rays[:, 4] = py.reshape(-1)
    rays[:, 5] = -1.0
    rays[:, 7] = INFTY
    return rays


def incoherent_rays(n, lo, hi, seed=2, start=0, anyhit=False):
    """Origins uniform in the AABB [lo,hi] inflated by 10 %, directions uniform on the sphere,
    from a counter-based generator keyed (ray index, seed). float32 [n,8].

    anyhit=True sets tmax to the distance to a second uniform point (shadow-ray style)."""
    lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
    c = 0.5 * (lo + hi); e = 0.5 * (hi - lo) * 1.1
    idx = np.arange(start, start + n, dtype=np.uint64)
    rays = np.zeros((n, 8), dtype=np.float32)
    for k in range(3):
        rays[:, k] = (c[k] + e[k] * (2.0 * _u01(idx, k, seed) - 1.0)).astype(np.float32)
    z = 2.0 * _u01(idx, 3, seed) - 1.0
    phi = 2.0 * np.pi * _u01(idx, 4, seed)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    rays[:, 3] = (s * np.cos(phi)).astype(np.float32)
    rays[:, 4] = (s * np.sin(phi)).astype(np.float32)
    rays[:, 5] = z.astype(np.float32)
    if anyhit:
        q = np.stack([c[k] + e[k] * (2.0 * _u01(idx, 5 + k, seed) - 1.0) for k in range(3)], -1)
        d = q - rays[:, 0:3].astype(np.float64)
        rays[:, 3:6] = d.astype(np.float32)
        rays[:, 7] = np.sqrt((d * d).sum(-1)).astype(np.float32)
    else:
        rays[:, 7] = INFTY
    return rays


def cube_triangles():
    """Unit cube [0,1]^3 as 12 triangles (same shape as the reference fixture box.ply: 8 v / 12 f)."""
    v = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, c, d in quads:
        f += [(a, b, c), (a, c, d)]
    return v, np.array(f, dtype=np.int32)
