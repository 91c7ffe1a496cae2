"""Development tool: host-buffer ray-cast call (spb_trace_closest) at several chunk sizes on the C2 workload."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

n = 1 << 24
v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
rays = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2)
ctx = capi.Context(0)
ctx.set_triangles(tris)
ctx.build(max_leaf_tris=1)
pin_rays = torch.from_numpy(rays).pin_memory().numpy()
pin_hits = torch.empty(n * 16, dtype=torch.uint8).pin_memory().numpy().view(capi.HIT)
for chunk in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["262144", "524288", "1048576", "2097152"])]:
    ctx.set_option("chunk_rays", chunk)
    for _ in range(2):
        ctx.trace_closest(pin_rays, out=pin_hits)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        ctx.trace_closest(pin_rays, out=pin_hits)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"chunk": chunk, "best_mrays": n / min(ts) * 1e-6, "median_mrays": n / float(np.median(ts)) * 1e-6}), flush=True)
