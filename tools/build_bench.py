"""Builder comparison (development tool): python tools/build_bench.py [nu nv] ...
For each torus(nu, nv): build with the device SAH builder (0), the host SAH builder (2) and the device LBVH (1); build time,
tree statistics, byte equality of 0 vs 2, and the C2-style incoherent closest-hit rate on each tree."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(1000, 500)]
n_rays = int(os.environ.get("SPB_RAYS", str(1 << 24)))
out = []
for nu, nv in sizes:
    v, f = scenes.torus_mesh(nu, nv)
    tris = scenes.mesh_triangles(v, f)
    rays = scenes.incoherent_rays(n_rays, v.min(0), v.max(0), seed=2)
    ctx = capi.Context(0)
    ctx.set_triangles(tris)
    d_rays = ctx.dev_alloc(n_rays * 32); d_hits = ctx.dev_alloc(n_rays * 16)
    ctx.dev_upload(d_rays, rays)
    blobs = {}
    ref_hits = None
    for builder in ([0, 0, 2, 1] if len(tris) <= 2_000_000 else [0, 0, 2]):
        t0 = time.perf_counter()
        ctx.build(builder=builder)
        wall = time.perf_counter() - t0
        st = ctx.stats()
        ms = []
        for _ in range(4):
            ctx.trace_closest_dev(d_rays, n_rays, d_hits)
            ms.append(ctx.counters()["last_kernel_ms"])
        hits = np.empty(n_rays, dtype=capi.HIT); ctx.dev_download(hits, d_hits)
        if ref_hits is None:
            ref_hits = hits.copy()
        r = {"triangles": len(tris), "builder": builder, "build_wall_s": wall, "build_s": st["build_seconds"], "wide_nodes": st["n_wide_nodes"],
             "max_depth": st["max_depth"], "sah_cost": st["sah_cost"], "mrays_s": n_rays / min(ms[1:]) * 1e-3,
             "same_prims_as_first": bool(np.array_equal(hits["prim"], ref_hits["prim"]))}
        if len(tris) <= 2_000_000:
            blobs[builder] = tuple(bytes(x) for x in ctx.export_bvh()[1:])
            if builder == 2:
                r["bytes_equal_device_vs_host"] = blobs[0] == blobs[2]
        print(json.dumps(r), flush=True)
        out.append(r)
    ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/build_bench.json", "w"), indent=1)
