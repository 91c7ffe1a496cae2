#!/bin/bash
# 8-GPU session (r01g kernels): render configs C3/C4/C5 at 8 GPUs (full spp) through the host CLI, then the ray-cast bench under torchrun at N=8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python tools/render_configs.py 8 c3,c4 1.0 2>&1 | tail -2 | cut -c1-420
cp gpurun_out/render_configs_g8.json gpurun_out/render_configs_g8_c3c4.json
timeout 400 python tools/render_configs.py 8 c5 1.0 2>&1 | tail -1 | cut -c1-600
cp gpurun_out/render_configs_g8.json gpurun_out/render_configs_g8_c5.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_n8_err.txt | tail -1 | tee gpurun_out/bench_n8.json | cut -c1-300
tail -3 gpurun_out/bench_n8_err.txt
