#!/bin/bash
# r02ab (2 GPUs): cross-process film sum over CUDA IPC: tests, bench N=2 with ipc and with nccl; CLI start-up with the device mask
mkdir -p gpurun_out /tmp/rc
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_host.py -x -q -m gpu 2>&1 | grep -v "^\[INFO\]" | tail -n 12 | tee gpurun_out/r02ab_pytest.txt
for how in ipc nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --film-reduce $how 2> gpurun_out/r02ab_bench_n2_${how}_err.txt > gpurun_out/r02ab_bench_n2_$how.json
  tail -n 3 gpurun_out/r02ab_bench_n2_${how}_err.txt | cut -c1-300
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02ab_bench_n2_$how.json').read())
for k in ('render_c3','render_c4'):
    r=d['extra'][k]; print('$how', k, round(r['msamples_s'],1), 'sec', round(r['seconds'],5), 'allreduce_ms', round(r['allreduce_ms'],3), 'render_ms', round(r['render_ms_slowest_rank'],2), r['allreduce'][:40])
PY
done 2>&1 | tee gpurun_out/r02ab_bench_summary.txt
python -c "
from spica_b200 import scenes
scenes.write_cornell('/tmp/rc', 1920, 1080, 256, 16, variant='diffuse', name='c3')
" 2>&1 | tail -n 2
export SPICA_TIMING=1
( cd spica_b200/bin
for mode in masked all masked all; do
  echo "== c3 gpus 1 ($mode devices visible)"
  if [ $mode = all ]; then export SPICA_ALL_DEVICES_VISIBLE=1; else unset SPICA_ALL_DEVICES_VISIBLE; fi
  ./spica -i /tmp/rc/c3.xml -o /tmp/rc/c3_out --gpus 1 --seed 1 2>&1 | grep -E "TIME|rendered|rror" | grep -E "context of GPU 0|parse\(\) returned|rendered|rror"
done ) 2>&1 | tee gpurun_out/r02ab_cli_device_mask.txt
