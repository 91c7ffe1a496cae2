#!/bin/bash
# GPU session 4 (1 GPU): parity tests incl. the reference-host plugin shim, smoke, the reference CLI on Cornell C1 with the GPU plugins.
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== reference host + GPU plugins, Cornell C1 (512x512, 64 spp, depth 8)"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/refhost_c1.txt
import os, time
from spica_b200 import refhost, scenes
REF = os.path.join(os.getcwd(), "oracle", "_ref")
xml = scenes.write_cornell("/tmp/c1", 512, 512, 64, 8, variant="diffuse", name="c1")
for plugins in [("path", "bvh"), ("path",)]:
    t0 = time.time()
    r = refhost.run(xml, "/tmp/c1/out", "/tmp/c1/run", REF, env={"SPICA_SEED": 1}, gpu_plugins=plugins)
    dt = time.time() - t0
    print("plugins swapped:", plugins, "rc", r.returncode, "wall %.2f s (process start to exit)" % dt)
    print("\n".join(l for l in r.stdout.splitlines() if "INFO" in l or "Finish" in l))
    print(r.stderr[-500:])
    print("mean radiance", scenes.read_hdr("/tmp/c1/out.hdr").mean())
PY
ls -la gpurun_out
