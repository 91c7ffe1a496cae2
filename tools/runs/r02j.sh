#!/bin/bash
# r02j (8 GPUs): bench.py under torchrun at N=8: C2 weak scaling, C3 / C4 / C5 strong-scaled frames with the NCCL film reduce
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r02j_bench_n8_err.txt > gpurun_out/r02j_bench_n8.json ) 2>&1 | tail -n 4
cut -c1-300 gpurun_out/r02j_bench_n8.json; grep -v "^\[W\|^W1\|^\*\*\*\|spb build" gpurun_out/r02j_bench_n8_err.txt | tail -n 20
