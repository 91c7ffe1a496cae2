#!/bin/bash
# r02u (1 GPU): queue capacity at small and large calls (32 spp = one GPU's share of C4 on 8 GPUs; 128 spp = of C3)
mkdir -p gpurun_out
timeout 300 python tools/render_bench.py diffuse,glossy 32 4194304,8388608,16777216 1 2>&1 | tee gpurun_out/r02u_render_bench_32spp.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608,16777216,33554432 1 2>&1 | tee gpurun_out/r02u_render_bench_128spp.txt
