#!/bin/bash
# GPU session 6 (1 GPU): re-verify HEAD on a fresh box: parity tests, smoke, bench (both arms).
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.txt
