"""Debug: where do the device-built and host-built wide BVHs differ?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes

NODE = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"), ("tri_base", "<u4"), ("meta", "u1", 8),
                 ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
TRI = np.dtype([("v0", "<f4", 3), ("id", "<i4"), ("v1", "<f4", 3), ("rank", "<i4"), ("v2", "<f4", 3), ("pad", "<i4")])
assert NODE.itemsize == 80 and TRI.itemsize == 48
g = dict(np.load("tests/golden/raycast_torus_1600.npz"))
tris = g["verts"].astype(np.float64)[g["faces"]].reshape(-1, 9)
ctx = capi.Context(0)
ctx.set_triangles(tris)
for ml in (1, 3):
    ctx.build(max_leaf_tris=ml, builder=2); sh, nh, th = ctx.export_bvh()
    ctx.build(max_leaf_tris=ml, builder=0); sd, nd, td = ctx.export_bvh()
    nh = nh.view(NODE); nd = nd.view(NODE); th = th.view(TRI); td = td.view(TRI)
    print("max_leaf", ml, "nodes", len(nh), len(nd), "tris", len(th), len(td))
    dn = [i for i in range(min(len(nh), len(nd))) if nh[i].tobytes() != nd[i].tobytes()]
    dt = np.nonzero(th["id"] != td["id"])[0]
    print(" differing nodes", len(dn), dn[:10], " differing tri ids", len(dt), dt[:10])
    print(" same tri id multiset", np.array_equal(np.sort(th["id"]), np.sort(td["id"])))
    for i in dn[:3]:
        print("  node", i, "\n   host", nh[i], "\n   dev ", nd[i])
    if len(dt):
        print("  host ids", th["id"][:16], "\n  dev  ids", td["id"][:16])
    # same set of nodes modulo numbering?  compare per-level sorted (p, e, qlo, qhi) signatures
    sig = lambda a: sorted((x["p"].tobytes() + x["e"].tobytes() + x["qlo"].tobytes() + x["qhi"].tobytes()) for x in a)
    print(" node box signatures equal as multisets:", sig(nh) == sig(nd))
