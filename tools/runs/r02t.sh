#!/bin/bash
# r02t (2 GPUs): the multi-GPU paths of both hosts on the final build (new refplugin two-GPU test), full GPU suite on a 2-GPU box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee gpurun_out/r02t_pytest_gpu_2gpus.txt
