#!/bin/bash
# r02i (1 GPU): cp.async-staged shade kernel + device-side shading records: tests, render A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02i_pytest_gpu.txt
SPICA_B200_LIB=$PWD/spica_b200/lib_ab/libspica_b200.so timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608,16777216 1 2>&1 | tee gpurun_out/r02i_render_before.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608,16777216 1 2>&1 | tee gpurun_out/r02i_render_staged.txt
