"""Render-loop throughput through the C ABI (development tool):
   python tools/render_bench.py [variants=diffuse,glossy] [spp=128] [slots=8388608,...] [graph=1,0]
Each configuration: one warm-up call, then the best of two timed spb_render_samples calls at 1920x1080, depth 16."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi  # noqa: E402

variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["diffuse", "glossy"]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
slots = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 23]
graphs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
W, H = 1920, 1080
out = []
for variant in variants:
    for sl in slots:
        for g in graphs:
            ctx = capi.Context(0)
            ctx.set_option("wave_slots", sl)
            ctx.set_option("render_graph", g)
            if os.environ.get("SPB_SHADE_MINB"):
                ctx.set_option("shade_minb", int(os.environ["SPB_SHADE_MINB"]))
            img = capi.cornell_render(ctx, W, H, 8, max_depth=16, seed=1, variant=variant)
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                ctx.render_samples(8 + spp * rep, spp, 1)
                best = min(best, time.perf_counter() - t0)
            img = ctx.film_resolve()
            st = ctx.render_stats()
            r = {"scene": variant, "slots": sl, "graph": g, "spp": spp, "shade_minb": os.environ.get("SPB_SHADE_MINB", "default"), "msamples_s": W * H * spp / best * 1e-6, "seconds": best,
                 "mean": float(img.mean()), "iterations": st["iterations"], "launches": st["kernel_launches"],
                 "rays_per_sample": (st["rays_closest"] + st["rays_shadow"] + st["rays_mis"]) / max(st["paths"], 1)}
            print(json.dumps(r), flush=True)
            out.append(r)
            ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/render_bench.json", "w"), indent=1)
