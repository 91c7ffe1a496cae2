#!/bin/bash
# r02a (1 GPU): A/B of the node test (I2F on the XU pipe vs PRMT on the ALU pipe) on the C2 workload, trace parity
# tests, and ncu --set full captures of the shade kernels and of the traversal kernels inside a render.
mkdir -p gpurun_out
SPICA_B200_LIB=$PWD/spica_b200/lib_ab/libspica_b200.so timeout 300 python tools/sweep4.py 16777216 5 render > gpurun_out/r02a_sweep_i2f.txt 2>&1
timeout 300 python tools/sweep4.py 16777216 5 render > gpurun_out/r02a_sweep_prmt.txt 2>&1
tail -n 5 gpurun_out/r02a_sweep_i2f.txt gpurun_out/r02a_sweep_prmt.txt
timeout 600 python -m pytest tests/test_trace_gpu.py -x -q 2>&1 | tail -n 3 | tee gpurun_out/r02a_pytest_trace.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -c 2 -o gpurun_out/r02a_shade_diffuse python tools/render_once.py diffuse 4 > gpurun_out/r02a_ncu_shade_diffuse.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -c 2 -o gpurun_out/r02a_shade_glossy python tools/render_once.py glossy 4 > gpurun_out/r02a_ncu_shade_glossy.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traceCoop -c 3 -o gpurun_out/r02a_trace_in_render python tools/render_once.py diffuse 4 > gpurun_out/r02a_ncu_trace_render.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traceCoopPair -s 2 -c 1 -o gpurun_out/r02a_trace_c2 python tools/sweep4.py 16777216 5 > gpurun_out/r02a_ncu_trace_c2.log 2>&1
bash tools/ncu_to_csv.sh gpurun_out/*.ncu-rep
du -sh gpurun_out
