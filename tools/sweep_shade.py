"""Development tool: Cornell renders with the all-lobes shade kernel and with the scene-specialised instances
(and the diffuse instance compiled for 4 / 5 / 6 resident CTAs per SM); every image must equal the generic one bit for bit."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi  # noqa: E402

out = []
for variant, configs in (("diffuse", [(1, 0), (0, 4), (0, 5), (0, 6)]), ("glossy", [(1, 0), (0, 0)])):
    ref = None
    for generic, mb in configs:
        c2 = capi.Context(0)
        c2.set_option("shade_generic", generic)
        c2.set_option("shade_minb", mb)
        img = capi.cornell_render(c2, 1920, 1080, 16, max_depth=16, seed=1, variant=variant)   # warm-up
        dt = 1e9
        for rep in range(2):
            t0 = time.perf_counter()
            img = capi.cornell_render(c2, 1920, 1080, 128, max_depth=16, seed=1, variant=variant, first=16 + 128 * rep, begin=False)
            dt = min(dt, time.perf_counter() - t0)
        if ref is None:
            ref = img.copy()
        r = {"scene": variant, "generic": generic, "shade_minb": mb, "msamples_s": 1920 * 1080 * 128 / dt * 1e-6,
             "mean": float(img.mean()), "identical_to_generic": bool(np.array_equal(img, ref))}
        print(json.dumps(r), flush=True)
        out.append(r)
        c2.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_shade.json", "w"), indent=1)
