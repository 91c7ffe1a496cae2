#!/bin/bash
# r02k (2 GPUs): the hosts' multi-GPU path (clone + reduce to GPU 0), the refplugin with SPICA_GPUS=2, bench N=2 (reduce_ms after the watchdog change)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host.py tests/test_refplugin.py -x -q -m gpu 2>&1 | grep -v "^\[INFO\]" | tail -n 12 | tee gpurun_out/r02k_pytest_hosts.txt
timeout 600 python tools/render_configs.py 2 c3,c4 0.25 2>&1 | tail -n 4 | tee gpurun_out/r02k_render_configs_g2.txt
timeout 600 python tools/render_configs.py 1 c3,c4 0.25 2>&1 | tail -n 4 | tee gpurun_out/r02k_render_configs_g1.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r02k_bench_n2_err.txt > gpurun_out/r02k_bench_n2.json ) 2>&1 | tail -n 4
wc -l gpurun_out/r02k_bench_n2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_n2.json').read())
for k in ('render_c3','render_c4'):
    r=d['extra'][k]; print(k, round(r['msamples_s'],1), 'allreduce_ms', round(r['allreduce_ms'],3))
print('e2e', d['e2e'])
PY
