"""GPU tuning sweep (development tool): Mrays/s of the trace kernels over builder / launch options."""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
rays = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2)
pri = scenes.primary_rays(2048, 2048)[:n]
ctx = capi.Context(0)
ctx.set_triangles(tris)
d_rays = ctx.dev_alloc(n * 32)
d_pri = ctx.dev_alloc(len(pri) * 32)
d_hits = ctx.dev_alloc(n * 16)
ctx.dev_upload(d_rays, rays)
ctx.dev_upload(d_pri, pri)
res = []
for max_leaf in (1, 2, 3):
    ctx.build(max_leaf_tris=max_leaf)
    st = ctx.stats()
    ctx.set_option("counters", 1)
    c0 = ctx.counters()
    ctx.trace_closest_dev(d_rays, n, d_hits)
    c1 = ctx.counters()
    ctx.set_option("counters", 0)
    nv = (c1["node_visits"] - c0["node_visits"]) / n
    tt = (c1["tri_tests"] - c0["tri_tests"]) / n
    for variant, cps in [(0, 0), (1, 0), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)]:
        ctx.set_option("trace_variant", variant)
        ctx.set_option("trace_ctas_per_sm", cps)
        ms = []
        for it in range(4):
            ctx.trace_closest_dev(d_rays, n, d_hits)
            ms.append(ctx.counters()["last_kernel_ms"])
        ctx.trace_closest_dev(d_pri, len(pri), d_hits)
        pms = ctx.counters()["last_kernel_ms"]
        r = {"max_leaf": max_leaf, "variant": variant, "ctas_per_sm": cps, "inc_mrays": n / min(ms[1:]) * 1e-3,
             "pri_mrays": len(pri) / pms * 1e-3, "nodes_per_ray": nv, "tris_per_ray": tt,
             "wide_nodes": st["n_wide_nodes"], "build_s": st["build_seconds"]}
        print(json.dumps(r), flush=True)
        res.append(r)
json.dump(res, open(os.path.join("gpurun_out", "sweep.json"), "w"), indent=1)
