"""Measures V_int / V_leaf: mean inner-node and leaf visits per ray of an ordered,
tmax-pruned traversal of the REFERENCE-built binary tree, for BASELINE config 2 (1M-triangle torus;
incoherent and primary ray sets).  These constants define the algorithmic bytes per ray used by
bench.py's roofline.  Run from the repo root: python tests/golden/make_visits.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from spica_b200 import scenes  # noqa: E402

v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
nodes = ob.bvh_build(tris)
n = 1 << 20
inc = scenes.incoherent_rays(1 << 24, v.min(0), v.max(0), seed=2)[:: (1 << 24) // n]
pri = scenes.primary_rays(4096, 4096)[:: (1 << 24) // n]
anyr = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2, anyhit=True)
out = {"scene": "torus(nu=1000,nv=500) 1,000,000 triangles", "sample_rays": n,
       "definition": "ordered near-first traversal of the reference-built binary BVH, children pruned by the shrinking tmax"}
for name, rays in (("incoherent", inc), ("primary", pri), ("anyhit", anyr)):
    vi, vl = ob.count_ordered_visits(nodes, tris, rays)
    p, t, _, _ = ob.trace_closest(nodes, tris, rays)
    out[name] = {"V_int": vi, "V_leaf": vl, "hit_fraction": float((p >= 0).mean()),
                 "bytes_per_ray_f32verts": 32 + 16 + vi * 64 + vl * 48}
    print(name, out[name])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "visits_c2.json"), "w"), indent=1)
