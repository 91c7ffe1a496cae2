#!/bin/bash
# ncu --set full capture of one closest-hit launch for each given kernel variant (development tool)
mkdir -p gpurun_out
for v in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace\(Persistent\|Coop\) -s 2 -c 1 -f -o gpurun_out/prof_v$v \
      python tools/sweep2.py 8388608 $v > gpurun_out/ncu_v$v.log 2>&1
  tail -2 gpurun_out/ncu_v$v.log
done
