#!/bin/bash
# r02h (1 GPU): A/B of the batched ray cursor (lib_ab = one atomic per refill), ncu of the shade kernel after the CTA-aggregated pushes
mkdir -p gpurun_out
SPICA_B200_LIB=$PWD/spica_b200/lib_ab/libspica_b200.so timeout 300 python tools/sweep4.py 16777216 5 > gpurun_out/r02h_sweep_before.txt 2>&1
timeout 300 python tools/sweep4.py 16777216 5 > gpurun_out/r02h_sweep_batched.txt 2>&1
tail -n 2 gpurun_out/r02h_sweep_before.txt gpurun_out/r02h_sweep_batched.txt
SPICA_B200_LIB=$PWD/spica_b200/lib_ab/libspica_b200.so timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608 1 2>&1 | tee gpurun_out/r02h_render_before.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608 1 2>&1 | tee gpurun_out/r02h_render_batched.txt
timeout 300 python -m pytest tests/test_trace_gpu.py -x -q 2>&1 | tail -n 3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 6 -c 1 -o gpurun_out/r02h_shade_diffuse python tools/render_once.py diffuse 16 > gpurun_out/r02h_ncu_shade_diffuse.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 18 -c 3 -o gpurun_out/r02h_shade_glossy python tools/render_once.py glossy 16 > gpurun_out/r02h_ncu_shade_glossy.log 2>&1
bash tools/ncu_to_csv.sh gpurun_out/*.ncu-rep
