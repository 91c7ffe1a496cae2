"""First build of a process against the following ones (development tool): python tools/build_cold.py [nu=2500] [nv=2000] [builders=0,0,0]
SPICA_BUILD_TIMING=1 prints the device builder's phases on stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SPICA_BUILD_TIMING"] = "1"
from spica_b200 import capi, scenes
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
builders = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 0, 0]
v, f = scenes.torus_mesh(nu, nv)
tris = scenes.mesh_triangles(v, f)
t0 = time.perf_counter(); ctx = capi.Context(0); print("context %.3f s" % (time.perf_counter() - t0), flush=True)
t0 = time.perf_counter(); ctx.set_triangles(tris); print("set_triangles %.3f s" % (time.perf_counter() - t0), flush=True)
for rep, b in enumerate(builders):
    t0 = time.perf_counter(); ctx.build(builder=b); print("build %d (builder %d): %.3f s wall, build_seconds %.3f" % (rep, b, time.perf_counter() - t0, ctx.stats()["build_seconds"]), flush=True)
ctx.close()
