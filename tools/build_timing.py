"""Where does the device builder's wall time go, and what does order of use do to it?  (development tool)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SPICA_BUILD_TIMING"] = "1"
import torch
from spica_b200 import capi, scenes
torch.cuda.set_device(0)
v, f = scenes.torus_mesh(1000, 500)
tris = scenes.mesh_triangles(v, f)
ctx = capi.Context(0)
ctx.set_triangles(tris)
for label, b in (("host", 2), ("device", 0), ("device", 0), ("host", 2), ("device", 0)):
    t0 = time.perf_counter(); ctx.build(builder=b); print(label, "%.3f s wall, build_seconds %.3f" % (time.perf_counter() - t0, ctx.stats()["build_seconds"]), flush=True)
    st, nodes, tr = ctx.export_bvh()
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); del x
t0 = time.perf_counter(); ctx.build(); print("device after a torch alloc/free %.3f s" % (time.perf_counter() - t0), flush=True)
# many small contexts, like bench.py's C1 runs
for rep in range(4):
    t0 = time.perf_counter()
    c = capi.Context(0)
    img = capi.cornell_render(c, 512, 512, 64, max_depth=8, seed=rep)
    t1 = time.perf_counter()
    c.close()
    print("C1 through capi: render %.3f s, close %.3f s" % (t1 - t0, time.perf_counter() - t1), flush=True)
