"""Image-parity fixtures AT THE BASELINE CONFIGURATIONS, from the UNMODIFIED reference renderer (oracle/_ref/bin/spica):

  c1 : BASELINE.json configs[0]  Cornell box 512x512, 64 spp, max depth 8          -> baseline_c1_ref.npz
  c4 : BASELINE.json configs[3]  glossy Cornell box 1920x1080, max depth 16, at a REDUCED 16 spp (every pass is
       identical work, core/integrator.cc:64; 256 spp would take the CPU reference half an hour per run)
                                                                                    -> baseline_c4_16spp_ref.npz

K = 4 reference runs each (the reference seeds from time(0), core/integrator.cc:51,71).  To keep the fixtures small the
images are stored box-averaged over bin x bin pixel blocks (c1: 2, c4: 4); tests/test_render_gpu.py bins the GPU image
the same way, so the relMSE threshold (1.5 x the largest pairwise relMSE of the binned reference runs) is like for like.

  python tests/golden/make_baseline_golden.py [c1] [c4]
"""
import itertools
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spica_b200 import scenes  # noqa: E402

K = 4
# c5 : BASELINE.json configs[4] IN MINIATURE -- the same scene (torus + ground under the environment map, rough dielectric, tent
#      filter, depth 16) with 100,000 instead of 10,000,000 triangles (the torus is smooth: the tessellation only changes
#      facet size) at 480x270, 64 spp; the full 10 M-triangle scene needs 10 GB and minutes of build time in the reference
#                                                                                   -> baseline_c5_small_ref.npz
CONFIGS = {"c5": dict(w=480, h=270, spp=64, depth=16, variant="env", bin=2, out="baseline_c5_small_ref.npz", nu=250, nv=200),"c1": dict(w=512, h=512, spp=64, depth=8, variant="diffuse", bin=2, out="baseline_c1_ref.npz"),
           "c4": dict(w=1920, h=1080, spp=16, depth=16, variant="glossy", bin=4, out="baseline_c4_16spp_ref.npz")}


def main():
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin")
    for name in sys.argv[1:] or ["c1", "c4"]:
        c = CONFIGS[name]
        if c["variant"] == "env":
            xml = scenes.write_envscene("/tmp/spica_baseline_golden", c["w"], c["h"], c["spp"], c["depth"], name=name, nu=c["nu"], nv=c["nv"])
        else:
            xml = scenes.write_cornell("/tmp/spica_baseline_golden", c["w"], c["h"], c["spp"], c["depth"], variant=c["variant"], name=name)
        runs, secs = [], []
        for k in range(K):
            out = "/tmp/spica_baseline_golden/%s_%d" % (name, k)
            t0 = time.time()
            subprocess.run(["./spica", "-i", xml, "-t", str(os.cpu_count() or 8), "-o", out], cwd=ref_bin, check=True, stdout=subprocess.DEVNULL)
            secs.append(time.time() - t0)
            runs.append(scenes.bin_image(scenes.read_hdr(out + ".hdr"), c["bin"]))
            print(name, "run", k, "%.1fs" % secs[-1], "mean", runs[-1].mean(), flush=True)
            time.sleep(1.2)
        runs = np.stack(runs)
        m = runs.mean(0)
        pair = [scenes.rel_mse(runs[i], runs[j], m) for i, j in itertools.combinations(range(K), 2)]
        print(name, "pairwise relMSE (binned) max %.6f mean %.6f" % (max(pair), np.mean(pair)))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", c["out"]), runs=runs.astype(np.float16), spp=c["spp"], width=c["w"],
                            height=c["h"], max_depth=c["depth"], bin=c["bin"], nu=c.get("nu", 0), nv=c.get("nv", 0), pair_relmse=np.array([max(pair), np.mean(pair)]),
                            cpu_seconds=np.array(secs), cpu_threads=os.cpu_count() or 8)


if __name__ == "__main__":
    main()
