#!/bin/bash
# r02b (1 GPU): the streaming integrator (regeneration, queue-indexed state, per-bucket shading, CUDA graph): parity tests,
# smoke, render throughput for queue sizes / graph on-off, C2 trace sweep with the PRMT node test.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02b_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 | tee gpurun_out/r02b_smoke.txt
timeout 600 python tools/render_bench.py diffuse,glossy 128 2097152,4194304,8388608,16777216 1 2>&1 | tee gpurun_out/r02b_render_bench.txt
timeout 300 python tools/render_bench.py diffuse,glossy 128 8388608 0 2>&1 | tee -a gpurun_out/r02b_render_bench.txt
timeout 300 python tools/sweep4.py 16777216 5 2>&1 | tee gpurun_out/r02b_sweep.txt
