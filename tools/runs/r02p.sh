#!/bin/bash
# r02p (8 GPUs): the C++ host CLI on 8 GPUs vs 1 GPU at the full BASELINE sample counts: wall clock from process start to image file
mkdir -p gpurun_out
export SPICA_TIMING=1
timeout 900 python tools/render_configs.py 8 c3,c4,c5 2>&1 | tail -n 3 | cut -c1-420 | tee gpurun_out/r02p_render_configs_g8.txt
cp gpurun_out/render_configs_g8.json gpurun_out/r02p_render_configs_g8.json
timeout 900 python tools/render_configs.py 1 c3,c4,c5 2>&1 | tail -n 3 | cut -c1-420 | tee gpurun_out/r02p_render_configs_g1.txt
cp gpurun_out/render_configs_g1.json gpurun_out/r02p_render_configs_g1.json
timeout 900 python tools/render_configs.py 8 c3,c4 2>&1 | tail -n 2 | cut -c1-420 | tee gpurun_out/r02p_render_configs_g8_again.txt
