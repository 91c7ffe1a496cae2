"""GPU tuning sweep 4 (development tool): traversal kernel variants on the C2 workload - closest hit on the
incoherent and primary ray sets and any-hit on the incoherent set; every variant's (prim, t, u, v) records
and occlusion flags are compared with the first variant's.  Optionally also a Cornell render per variant."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 5, 6]
render = len(sys.argv) > 3 and sys.argv[3] == "render"
v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
if os.environ.get("SPB_TRANSFORM"):     # the mesh under a double-precision toWorld rotation: vertices are no longer float32 numbers
    c, s_ = np.cos(0.7), np.sin(0.7)
    rot = np.array([[c, -s_, 0], [s_, c * np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    v = v.astype(np.float64) @ rot.T
    tris = np.ascontiguousarray(v[f].reshape(len(f), 9))
lo, hi = v.min(0), v.max(0)
rays = scenes.incoherent_rays(n, lo, hi, seed=2)
anyr = scenes.incoherent_rays(n, lo, hi, seed=2, anyhit=True)
pri = scenes.primary_rays(4096, 4096)[:n]
ctx = capi.Context(0)
ctx.set_triangles(tris)
ctx.build(max_leaf_tris=int(os.environ.get("SPB_MAX_LEAF", "0")))
print("tri_format", ctx.stats()["tri_format"], "lib", os.environ.get("SPICA_B200_LIB", "default"), flush=True)
d_rays = ctx.dev_alloc(n * 32); d_any = ctx.dev_alloc(n * 32); d_pri = ctx.dev_alloc(len(pri) * 32)
d_hits = ctx.dev_alloc(n * 16); d_occ = ctx.dev_alloc(n)
ctx.dev_upload(d_rays, rays); ctx.dev_upload(d_pri, pri); ctx.dev_upload(d_any, anyr)
ref = None
out = []


def best(fn, reps=4):
    ms = []
    for _ in range(reps):
        fn()
        ms.append(ctx.counters()["last_kernel_ms"])
    return min(ms[1:])


for var in variants:
    ctx.set_option("trace_variant", var)
    inc = best(lambda: ctx.trace_closest_dev(d_rays, n, d_hits))
    hits = np.empty(n, dtype=capi.HIT); ctx.dev_download(hits, d_hits)
    pm = best(lambda: ctx.trace_closest_dev(d_pri, len(pri), d_hits))
    phits = np.empty(len(pri), dtype=capi.HIT); ctx.dev_download(phits, d_hits)
    am = best(lambda: ctx.trace_any_dev(d_any, n, d_occ))
    occ = np.empty(n, dtype=np.uint8); ctx.dev_download(occ, d_occ)
    if ref is None:
        ref = (hits.copy(), phits.copy(), occ.copy())
    ok = bool(np.array_equal(hits, ref[0]) and np.array_equal(phits, ref[1]) and np.array_equal(occ, ref[2]))
    r = {"variant": var, "inc_mrays": n / inc * 1e-3, "pri_mrays": len(pri) / pm * 1e-3, "any_mrays": n / am * 1e-3, "same_hits": ok}
    print(json.dumps(r), flush=True)
    out.append(r)
if render:
    for var in variants:
        for variant in ("diffuse", "glossy"):
            c2 = capi.Context(0)
            c2.set_option("trace_variant", var)
            img = capi.cornell_render(c2, 1920, 1080, 16, max_depth=16, seed=1, variant=variant)   # warm-up: allocations + module load
            dt = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                img = capi.cornell_render(c2, 1920, 1080, 128, max_depth=16, seed=1, variant=variant, first=16 + 128 * rep, begin=False)
                dt = min(dt, time.perf_counter() - t0)
            r = {"variant": var, "scene": variant, "msamples_s": 1920 * 1080 * 128 / dt * 1e-6, "mean": float(img.mean())}
            print(json.dumps(r), flush=True)
            out.append(r)
            c2.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep4.json", "w"), indent=1)
