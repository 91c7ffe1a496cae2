#!/bin/bash
# r02g (2 GPUs): bench.py under torchrun at N=2 (NCCL film reduce, strong scaling of C3/C4), with builder phase timing on stderr
mkdir -p gpurun_out
export SPICA_BUILD_TIMING=1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r02g_bench_n2_err.txt > gpurun_out/r02g_bench_n2.json ) 2>&1 | tail -n 4
cut -c1-300 gpurun_out/r02g_bench_n2.json; grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02g_bench_n2_err.txt | tail -n 40
( time timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02g_bench_n1_err.txt > gpurun_out/r02g_bench_n1.json ) 2>&1 | tail -n 4
tail -n 30 gpurun_out/r02g_bench_n1_err.txt
