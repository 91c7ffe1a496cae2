"""Builder comparison on the C2 mesh (development tool): host SAH vs GPU LBVH - build time, tree size,
visits per ray, Mrays/s. Optionally at 10M triangles (argv[2] = 10)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
big = len(sys.argv) > 2 and sys.argv[2] == "10"
v, f = scenes.torus_mesh(2500, 2000) if big else scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
rays = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2)
ctx = capi.Context(0)
ctx.set_triangles(tris)
d_rays = ctx.dev_alloc(n * 32); d_hits = ctx.dev_alloc(n * 16)
ctx.dev_upload(d_rays, rays)
ref = None
for builder in (0, 1):
    t0 = time.time(); ctx.build(builder=builder); wall = time.time() - t0
    st = ctx.stats()
    ctx.set_option("counters", 1); c0 = ctx.counters(); ctx.trace_closest_dev(d_rays, n, d_hits); c1 = ctx.counters(); ctx.set_option("counters", 0)
    ms = []
    for it in range(4):
        ctx.trace_closest_dev(d_rays, n, d_hits); ms.append(ctx.counters()["last_kernel_ms"])
    hits = np.empty(n, dtype=capi.HIT); ctx.dev_download(hits, d_hits)
    if ref is None: ref = hits.copy()
    print(json.dumps({"triangles": len(tris), "builder": builder, "build_wall_s": wall, "build_s": st["build_seconds"], "wide_nodes": st["n_wide_nodes"],
                      "max_depth": st["max_depth"], "sah": st["sah_cost"], "nodes_per_ray": (c1["node_visits"] - c0["node_visits"]) / n,
                      "tris_per_ray": (c1["tri_tests"] - c0["tri_tests"]) / n, "mrays": n / min(ms[1:]) * 1e-3,
                      "same_hits": bool(np.array_equal(hits["prim"], ref["prim"]) and np.array_equal(hits["t"], ref["t"]))}), flush=True)
