#!/bin/bash
# GPU session 7 (1 GPU): ncu full capture of the default closest-hit kernel (variant 4) at the bench's ray count + GPU tests
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traceCoop -s 2 -c 1 -f -o gpurun_out/prof_v4_16M \
    python tools/sweep2.py 16777216 4 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -8 | tee gpurun_out/pytest_gpu.txt
