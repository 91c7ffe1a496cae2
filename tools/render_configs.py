"""BASELINE.json render configs through the C++ host CLI (`spica -i scene.xml --gpus G`), samples/s per config.
usage: python tools/render_configs.py G [c3,c4,c5] [spp_scale]     (development / measurement tool)
  c3: Cornell 1920x1080, 1024 spp, depth 16      c4: glossy Cornell 1920x1080, 256 spp
  c5: 10M-triangle torus + environment map, 3840x2160, 256 spp
spp_scale < 1 renders fewer samples (every pass is identical work)."""
import json
import os
import re
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spica_b200 import host, scenes  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c3", "c4", "c5"]
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = []
d = "/tmp/render_configs"
for c in which:
    t0 = time.time()
    if c == "c3":
        w, h, spp = 1920, 1080, max(1, int(1024 * scale))
        xml = scenes.write_cornell(d, w, h, spp, 16, variant="diffuse", name="c3")
    elif c == "c4":
        w, h, spp = 1920, 1080, max(1, int(256 * scale))
        xml = scenes.write_cornell(d, w, h, spp, 16, variant="glossy", name="c4")
    else:
        w, h, spp = 3840, 2160, max(1, int(256 * scale))
        xml = scenes.write_envscene(d, w, h, spp, 16, name="c5", nu=2500, nv=2000)
    t_scene = time.time() - t0
    t0 = time.time()
    r = host.run_cli(xml, os.path.join(d, c + "_out"), gpus=G, seed=1)
    wall = time.time() - t0
    info = [l for l in r.stdout.splitlines() if "rendered" in l or "BVH" in l]
    m = re.search(r"in ([0-9.]+) s: ([0-9.]+) Msamples/s", r.stdout)
    rec = {"config": c, "gpus": G, "width": w, "height": h, "spp": spp, "rc": r.returncode, "scene_write_s": round(t_scene, 2),
           "cli_wall_s": round(wall, 2), "render_s": float(m.group(1)) if m else None, "msamples_s": float(m.group(2)) if m else None,
           "log": info, "stderr": r.stderr[-300:]}
    if r.returncode == 0:
        img = scenes.read_hdr(os.path.join(d, c + "_out.hdr"))
        rec["mean_radiance"] = float(img.mean())
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/render_configs_g%d.json" % G, "w"), indent=1)
