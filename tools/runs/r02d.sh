#!/bin/bash
# r02d (1 GPU): builder byte equality after the leaf-order fix, compute-sanitizer over the hot path, ncu of the new shade kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_gpu.py -x -q -k "device or export" 2>&1 | grep -v "^\[INFO\]" | tail -n 12 | tee gpurun_out/r02d_pytest_builder.txt
bash tools/runs/r02_sanitize.sh
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 6 -c 1 -o gpurun_out/r02d_shade_diffuse python tools/render_once.py diffuse 16 > gpurun_out/r02d_ncu_shade_diffuse.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 18 -c 3 -o gpurun_out/r02d_shade_glossy python tools/render_once.py glossy 16 > gpurun_out/r02d_ncu_shade_glossy.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traceCoopPair -s 18 -c 2 -o gpurun_out/r02d_trace_steady python tools/render_once.py diffuse 16 > gpurun_out/r02d_ncu_trace_steady.log 2>&1
bash tools/ncu_to_csv.sh gpurun_out/*.ncu-rep
du -sh gpurun_out
