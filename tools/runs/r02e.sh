#!/bin/bash
# r02e (1 GPU): CTA-aggregated queue pushes; render tests, render bench, builder byte test, full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02e_pytest_gpu.txt
timeout 600 python tools/render_bench.py diffuse,glossy 128 4194304,8388608,16777216 1 2>&1 | tee gpurun_out/r02e_render_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02e_launches_diffuse.csv python tools/render_once.py diffuse 32 > gpurun_out/r02e_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02e_launches_glossy.csv python tools/render_once.py glossy 32 > gpurun_out/r02e_ncu_launches_glossy.log 2>&1
