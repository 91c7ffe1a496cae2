#!/bin/bash
# GPU session 11 (1 GPU, short): final binary: parity tests, smoke, bench N=1 (no reference arm, no ncu)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.txt
timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json | cut -c1-200
