"""Generates tests/golden/raycast_*.npz from the UNMODIFIED reference compiled under oracle/_ref
(`make -C oracle ref`, needs /root/reference).  Run from the repo root:

    python tests/golden/make_golden.py

Each fixture holds the inputs (triangles, rays - both regenerated deterministically by
spica_b200.scenes, stored anyway so the fixture is self-contained) and the reference's outputs:
closest-hit (prim, t), any-hit flags and the reference-built BVH (for the import path).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from spica_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def fixture(name, verts, faces, n_inc, prim_res, seed):
    tris = scenes.mesh_triangles(verts, faces)
    lo, hi = verts.min(0), verts.max(0)
    inc = scenes.incoherent_rays(n_inc, lo, hi, seed=seed)
    pri = scenes.primary_rays(prim_res, prim_res, eye=(0.0, 0.0, 4.0), fov_deg=40.0)
    rays = np.concatenate([inc, pri], 0)
    anyr = scenes.incoherent_rays(n_inc, lo, hi, seed=seed + 1, anyhit=True)
    # float64 rays: same set, perturbed so that the directions/origins are true doubles
    r64 = rays.astype(np.float64)
    r64[:, :6] *= (1.0 + 1e-9 * np.arange(1, 7))[None, :]
    info, prim, t, nodes = ob.ref_raycast(rays, tris=tris, dump_bvh=True)
    _, occ = ob.ref_raycast(anyr, mode="any", tris=tris)
    _, prim64, t64 = ob.ref_raycast(r64, tris=tris)
    np.savez_compressed(os.path.join(OUT, "raycast_%s.npz" % name), verts=verts, faces=faces, rays=rays,
                        prim=prim, t=t, any_rays=anyr, occluded=occ, rays64=r64, prim64=prim64, t64=t64,
                        bvh_nodes=nodes)
    print(name, "tris", len(tris), "rays", len(rays), "hit", int((prim >= 0).sum()), "occluded", int(occ.sum()),
          "nodes", len(nodes))


if __name__ == "__main__":
    assert ob.have_ref(), "build the reference first: make -C oracle ref"
    v, f = scenes.torus_mesh(40, 20)
    fixture("torus_1600", v, f, 6144, 48, seed=11)
    v, f = scenes.cube_triangles()
    fixture("cube_12", v * 2.0 - 1.0, f, 4096, 32, seed=21)
    # non-float32-representable vertices (exercises the 80-byte float64 triangle records)
    v, f = scenes.torus_mesh(24, 12)
    v64 = v.astype(np.float64) * (1.0 + 1e-10) + 1e-11
    tris = v64[f].reshape(-1, 9)
    lo, hi = v64.min(0), v64.max(0)
    rays = scenes.incoherent_rays(4096, lo, hi, seed=31)
    info, prim, t, nodes = ob.ref_raycast(rays, tris=tris, dump_bvh=True)
    np.savez_compressed(os.path.join(OUT, "raycast_torus_f64verts.npz"), tris=tris, rays=rays, prim=prim, t=t,
                        bvh_nodes=nodes)
    print("torus_f64verts", len(tris), int((prim >= 0).sum()))
