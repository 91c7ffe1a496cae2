#!/bin/bash
# r02v (1 GPU): MIS launch on a side stream + 32 Mi default queues: render tests, render bench at the new default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_host.py -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 6 | tee gpurun_out/r02v_pytest_render.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 33554432 1,0 2>&1 | tee gpurun_out/r02v_render_bench.txt
timeout 300 python tools/render_bench.py diffuse,glossy 32 33554432 1 2>&1 | tee -a gpurun_out/r02v_render_bench.txt
