"""Development tool (measurement build, make EXP=1): Cornell renders at several shade-kernel occupancies."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi  # noqa: E402

for variant in ("diffuse", "glossy"):
    for mb in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "3", "4", "5", "6"])]:
        c2 = capi.Context(0)
        c2.set_option("shade_minb", mb)
        img = capi.cornell_render(c2, 1920, 1080, 16, max_depth=16, seed=1, variant=variant)   # warm-up
        dt = 1e9
        for rep in range(2):
            t0 = time.perf_counter()
            img = capi.cornell_render(c2, 1920, 1080, 128, max_depth=16, seed=1, variant=variant, first=16 + 128 * rep, begin=False)
            dt = min(dt, time.perf_counter() - t0)
        print(json.dumps({"scene": variant, "shade_minb": mb, "msamples_s": 1920 * 1080 * 128 / dt * 1e-6, "mean": float(img.mean())}), flush=True)
        c2.close()
