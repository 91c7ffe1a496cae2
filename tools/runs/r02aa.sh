#!/bin/bash
# r02aa (8 GPUs): final build: bench.py under torchrun at N=8 (C5 on float32 records, 32 Mi queues), the C++ host on 8 GPUs
# with the films summed over peer memory (and with NCCL for comparison), CLI phases
mkdir -p gpurun_out /tmp/rc
n=8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r02aa_bench_n${n}_err.txt > gpurun_out/r02aa_bench_n$n.json
wc -l gpurun_out/r02aa_bench_n$n.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02aa_bench_n$n.json').read())
print('N=$n value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pcie frac', round(d['e2e']['pcie_ceiling']['frac'],3))
for k in ('render_c3','render_c4','render_c5'):
    r=d['extra'].get(k)
    if r: print(' ', k, round(r['msamples_s'],1), 'Msamples/s', round(r['seconds'],4), 's allreduce_ms', round(r['allreduce_ms'],3), 'render_ms', round(r['render_ms_slowest_rank'],1), r.get('triangle_records'))
PY
export SPICA_TIMING=1
timeout 900 python tools/render_configs.py 8 c3,c4,c5 2>&1 | tail -n 4 | tee gpurun_out/r02aa_render_configs_g8_peers.txt
cp gpurun_out/render_configs_g8.json gpurun_out/r02aa_render_configs_g8_peers.json
SPICA_FILM_REDUCE=nccl timeout 900 python tools/render_configs.py 8 c3,c4 2>&1 | tail -n 3 | tee gpurun_out/r02aa_render_configs_g8_nccl.txt
( cd spica_b200/bin
for mode in peers nccl peers nccl; do
  echo "== c3 gpus 8 ($mode)"
  export SPICA_FILM_REDUCE=$mode
  ./spica -i /tmp/render_configs/c3.xml -o /tmp/rc/c3_out --gpus 8 --seed 1 2>&1 | grep -E "TIME|rendered|rror|per-GPU"
done
unset SPICA_FILM_REDUCE
echo "== c3 gpus 1"; ./spica -i /tmp/render_configs/c3.xml -o /tmp/rc/c3_out --gpus 1 --seed 1 2>&1 | grep -E "TIME|rendered|rror"
echo "== c5 gpus 8 (peers)"; ./spica -i /tmp/render_configs/c5.xml -o /tmp/rc/c5_out --gpus 8 --seed 1 2>&1 | grep -E "TIME|rendered|rror|per-GPU|BVH" ) 2>&1 | tee gpurun_out/r02aa_cli_phases_g8.txt
