#!/bin/bash
# r02x (1 GPU): why is bench.py's C5 frame slower per GPU than the CLI's?  Same frame three ways at 32 spp on one GPU.
mkdir -p gpurun_out
timeout 600 python tools/render_env_bench.py 32 2500 2000 0 3 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02x_env_capi.txt
timeout 600 python tools/render_configs.py 1 c5 0.125 2>&1 | tail -n 2 | tee gpurun_out/r02x_env_cli.txt
timeout 900 python bench.py --steps 2 --warmup 3 --c3-spp 0 --c4-spp 0 --c5-spp 32 --no-cpu-baseline 2> gpurun_out/r02x_bench_err.txt > gpurun_out/r02x_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench.json').read())
r=d['extra']['render_c5']; print('bench c5', round(r['msamples_s'],1), r['seconds'], r['render_ms_slowest_rank'], r['rank0'])
PY
