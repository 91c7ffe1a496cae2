"""Generates the image-parity fixtures by running the UNMODIFIED reference renderer
(oracle/_ref/bin/spica, built by `make -C oracle ref`) K times on the Cornell scenes written by
spica_b200.scenes.write_cornell. Run in the build container (needs /root/reference to have been
compiled); the fixtures travel, the reference does not.

  python tests/golden/make_cornell_golden.py [variant ...]

Writes tests/golden/cornell_<variant>_ref.npz:
  runs   float16 [K, H, W, 3]   the K reference images (RGBE-decoded without truncation bias)
  spp, width, height, max_depth
  pair_relmse   max / mean pairwise relMSE between runs -> tau = 1.5 * max
The reference seeds its samplers from time(0) (core/integrator.cc:51,71), so runs differ by
construction; runs are spaced > 1 s apart.
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

K = 8
W = H = 128
SPP = 64
DEPTH = 8


def main():
    variants = sys.argv[1:] or ["diffuse", "glossy", "zoo", "envtorus"]
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin")
    for variant in variants:
        d = os.path.join(ROOT, "tests", "golden", "scenes")
        if variant == "envtorus":
            xml = scenes.write_envscene(d, W, H, SPP, DEPTH, name="envtorus")
        elif variant == "direct":       # the "zoo" scene (it has a mirror) under the directlighting integrator
            xml = scenes.write_cornell(d, W, H, SPP, DEPTH, variant="zoo", name="cornell_direct", integrator="directlighting")
        else:
            xml = scenes.write_cornell(d, W, H, SPP, DEPTH, variant=variant, name="cornell_" + variant)
        runs = []
        for k in range(K):
            out = "/tmp/spica_golden_%s_%d" % (variant, k)
            t0 = time.time()
            subprocess.run(["./spica", "-i", xml, "-t", str(os.cpu_count() or 8), "-o", out], cwd=ref_bin, check=True,
                           stdout=subprocess.DEVNULL)
            runs.append(scenes.read_hdr(out + ".hdr"))
            print(variant, "run", k, "%.1fs" % (time.time() - t0), "mean", runs[-1].mean(), flush=True)
            time.sleep(1.2)
        runs = np.stack(runs)
        m = runs.mean(0)
        pair = [scenes.rel_mse(runs[i], runs[j], m) for i, j in itertools.combinations(range(K), 2)]
        print(variant, "pairwise relMSE max %.5f mean %.5f" % (max(pair), np.mean(pair)))
        fname = "envtorus_ref.npz" if variant == "envtorus" else "cornell_%s_ref.npz" % variant
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", fname),
                            runs=runs.astype(np.float16), spp=SPP, width=W, height=H, max_depth=DEPTH,
                            pair_relmse=np.array([max(pair), np.mean(pair)]))


if __name__ == "__main__":
    main()
