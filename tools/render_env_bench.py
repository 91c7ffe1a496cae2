"""C5-style frame (torus + ground under an environment map, 3840x2160, depth 16) through the C ABI, no torch in the process
(development tool):  python tools/render_env_bench.py [spp=32] [nu=2500] [nv=2000] [slots=0] [calls=3]
One 1-spp warm-up call as bench.py does, then `calls` timed spb_render_samples calls of `spp` samples each."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
slots = int(sys.argv[4]) if len(sys.argv) > 4 else 0
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 3
W, H = 3840, 2160
ctx = capi.Context(0)
if slots:
    ctx.set_option("wave_slots", slots)
t0 = time.perf_counter()
capi.envscene_render(ctx, W, H, 1, max_depth=16, nu=nu, nv=nv, seed=7)
print("setup + 1 spp: %.2f s; build %.3f s" % (time.perf_counter() - t0, ctx.stats()["build_seconds"]), flush=True)
first = 1
for c in range(calls):
    t0 = time.perf_counter()
    ctx.render_samples(first, spp, 1)
    dt = time.perf_counter() - t0
    first += spp
    st = ctx.render_stats()
    rays = st["rays_closest"] + st["rays_shadow"] + st["rays_mis"]
    print(json.dumps({"call": c, "spp": spp, "slots": slots, "seconds": dt, "render_ms": st["render_ms"], "msamples_s": W * H * spp / dt * 1e-6,
                      "mrays_s": rays / st["render_ms"] * 1e-3, "rays_per_sample": rays / max(st["paths"], 1),
                      "iterations": st["iterations"], "launches": st["kernel_launches"]}), flush=True)
ctx.close()
