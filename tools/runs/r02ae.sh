#!/bin/bash
# r02ae (1 GPU): where the first build of a process spends its time (10 M triangles, then 1 M)
mkdir -p gpurun_out
( timeout 300 python tools/build_cold.py 2500 2000; timeout 300 python tools/build_cold.py 1000 500 ) 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02ae_build_cold.txt
