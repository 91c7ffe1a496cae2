#!/bin/bash
# r02s (8 GPUs): final bench.py under torchrun at N=8 and N=4 (the driver's SCALE run, as it will launch it)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r02s_bench_n${n}_err.txt > gpurun_out/r02s_bench_n$n.json
  wc -l gpurun_out/r02s_bench_n$n.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_n$n.json').read())
print('N=$n value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pcie frac', round(d['e2e']['pcie_ceiling']['frac'],3))
for k in ('render_c3','render_c4','render_c5'):
    r=d['extra'].get(k)
    if r: print(' ', k, round(r['msamples_s'],1), 'Msamples/s', round(r['seconds'],4), 's allreduce_ms', round(r['allreduce_ms'],3), 'render_ms', round(r['render_ms_slowest_rank'],1))
PY
done
