#!/bin/bash
# GPU session 2 (run with gpurun --gpus 2): tests, 1-GPU bench, 2-GPU bench + host CLI, ncu capture at the full ray count.
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench N=1"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench2_err.txt | tee gpurun_out/bench_n2.json
tail -5 gpurun_out/bench2_err.txt
echo "== host CLI 1 vs 2 GPUs"
python - <<'PY' 2>&1 | tail -12 | tee gpurun_out/host_cli.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from spica_b200 import host, scenes
d = "/tmp/cli_scene"; xml = scenes.write_cornell(d, 960, 540, 64, 16)
for g in (1, 2):
    t0 = time.time(); r = host.run_cli(xml, "/tmp/cli_out_g%d" % g, gpus=g, seed=5); dt = time.time() - t0
    print("gpus", g, "rc", r.returncode, "wall %.2fs" % dt, [l for l in r.stdout.splitlines() if "rendered" in l], r.stderr[-300:])
a = scenes.read_hdr("/tmp/cli_out_g1.hdr"); b = scenes.read_hdr("/tmp/cli_out_g2.hdr")
print("1-GPU vs 2-GPU image: max abs diff", float(np.abs(a - b).max()), "relMSE", scenes.rel_mse(a, b, a))
PY
echo "== ncu launches"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --render-spp 2 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full (closest-hit kernel at 16.7M rays)"
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePersistent -s 1 -c 1 -f -o gpurun_out/prof_trace \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --render-spp 0 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full (shade kernel)"
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 3 -c 1 -f -o gpurun_out/prof_shade \
    python bench.py --rays 1048576 --steps 1 --warmup 1 --no-cpu-baseline --render-spp 2 > gpurun_out/ncu_shade.log 2>&1
ls -la gpurun_out
