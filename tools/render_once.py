"""One Cornell render through the C ABI (development tool, used under ncu): python tools/render_once.py VARIANT SPP [W H DEPTH]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import capi  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "diffuse"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
h = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 16
ctx = capi.Context(0)
t0 = time.perf_counter()
img = capi.cornell_render(ctx, w, h, spp, max_depth=depth, variant=variant, seed=1)
print("render %s %dx%d %d spp: %.3f s, mean %.5f, stats %s" % (variant, w, h, spp, time.perf_counter() - t0, img.mean(), ctx.render_stats()))
ctx.close()
