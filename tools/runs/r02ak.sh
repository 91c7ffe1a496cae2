#!/bin/bash
# r02ak (2 GPUs): the final build under torchrun at N=2 (scratch pool / arena per process, IPC film sum), and on GPU 0 alone:
# shade_minb 4/5/6 and 64 Mi queues at the 32 Mi default's settings
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r02ak_bench_n2_err.txt > gpurun_out/r02ak_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ak_bench_n2.json').read())
print('N=2 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
for k in ('render_c3','render_c4'):
    r=d['extra'][k]; print(' ', k, round(r['msamples_s'],1), 'sec', round(r['seconds'],5), 'allreduce_ms', round(r['allreduce_ms'],3), r['allreduce'][:30])
PY
for mb in 4 5 6; do SPB_SHADE_MINB=$mb timeout 300 python tools/render_bench.py diffuse 128 33554432 1 2>&1 | grep -v "^\[INFO\]"; done | tee gpurun_out/r02ak_shade_minb.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 67108864 1 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02ak_slots_64mi.txt
