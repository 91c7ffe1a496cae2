#!/bin/bash
# r02ag (1 GPU): where the sporadic 0.2-0.6 s inside "collapse + emission" sits (finer laps); host build -> device build as bench.py does
mkdir -p gpurun_out
( timeout 300 python tools/build_cold.py 2500 2000 0,0,0,0; timeout 300 python tools/build_cold.py 1000 500 2,0,0,2,0 ) 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02ag_build_cold.txt
