#!/bin/bash
# r02al (8 GPUs): the final build under torchrun at N=8, as the driver launches it (arena + pool per process, MIS any-hit on C5, IPC film sum)
mkdir -p gpurun_out
n=8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r02al_bench_n${n}_err.txt > gpurun_out/r02al_bench_n$n.json
wc -l gpurun_out/r02al_bench_n$n.json; tail -n 3 gpurun_out/r02al_bench_n${n}_err.txt | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r02al_bench_n$n.json').read())
print('N=$n value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pcie frac', round(d['e2e']['pcie_ceiling']['frac'],3))
for k in ('render_c3','render_c4','render_c5'):
    r=d['extra'].get(k)
    if r: print(' ', k, round(r['msamples_s'],1), 'Msamples/s', round(r['seconds'],5), 's allreduce_ms', round(r['allreduce_ms'],3), 'render_ms', round(r['render_ms_slowest_rank'],2), r['allreduce'][:30])
PY
