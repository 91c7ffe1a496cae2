"""GPU tuning sweep 2 (development tool): kernel variants of the closest-hit traversal on the C2 workload;
each variant's hits are compared with variant 1's."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 20, 21, 22, 23, 24]
v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
rays = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2)
pri = scenes.primary_rays(4096, 4096)[:n]
ctx = capi.Context(0)
ctx.set_triangles(tris)
ctx.build()
d_rays = ctx.dev_alloc(n * 32); d_pri = ctx.dev_alloc(len(pri) * 32); d_hits = ctx.dev_alloc(n * 16)
ctx.dev_upload(d_rays, rays); ctx.dev_upload(d_pri, pri)
ref = None
out = []
for var in variants:
    ctx.set_option("trace_variant", var)
    ms = []
    for it in range(4):
        ctx.trace_closest_dev(d_rays, n, d_hits)
        ms.append(ctx.counters()["last_kernel_ms"])
    hits = np.empty(n, dtype=capi.HIT)
    ctx.dev_download(hits, d_hits)
    if ref is None:
        ref = hits.copy()
    ok = bool(np.array_equal(hits["prim"], ref["prim"]) and np.array_equal(hits["t"], ref["t"]))
    pm = []
    for it in range(3):
        ctx.trace_closest_dev(d_pri, len(pri), d_hits)
        pm.append(ctx.counters()["last_kernel_ms"])
    r = {"variant": var, "inc_mrays": n / min(ms[1:]) * 1e-3, "pri_mrays": len(pri) / min(pm[1:]) * 1e-3, "same_hits": ok}
    print(json.dumps(r), flush=True)
    out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep2.json", "w"), indent=1)
