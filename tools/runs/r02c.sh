#!/bin/bash
# r02c (1 GPU): the device SAH builder (byte equality with the host builder, edge cases, export / import / clone), builder
# timings at 1 M and 10 M triangles, the new render tests, and a launch list of a diffuse render
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_gpu.py -x -q -k "device or export or lbvh or collapse" 2>&1 | grep -v "^\[INFO\]" | tail -n 25 | tee gpurun_out/r02c_pytest_builder.txt
timeout 600 python tools/build_bench.py 1000 500 2>&1 | tee gpurun_out/r02c_build_bench.txt
SPB_RAYS=4194304 timeout 600 python tools/build_bench.py 2500 2000 2>&1 | tee -a gpurun_out/r02c_build_bench.txt
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_trace_gpu.py -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02c_pytest_render.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_launches_diffuse.csv python tools/render_once.py diffuse 32 > gpurun_out/r02c_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02c_launches_glossy.csv python tools/render_once.py glossy 32 > gpurun_out/r02c_ncu_launches_glossy.log 2>&1
du -sh gpurun_out
