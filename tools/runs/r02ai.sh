#!/bin/bash
# r02ai (1 GPU): per-kernel share of a C5 frame (10 M triangles + environment map, 3840x2160) -- ncu launch list at 4 spp
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02ai_launches_c5.csv python tools/render_env_bench.py 4 2500 2000 0 1 > gpurun_out/r02ai_c5_under_ncu.log 2>&1
tail -n 3 gpurun_out/r02ai_c5_under_ncu.log | cut -c1-300
python tools/launch_summary.py gpurun_out/r02ai_launches_c5.csv 2>&1 | head -n 40 | tee gpurun_out/r02ai_launch_share_c5.txt
