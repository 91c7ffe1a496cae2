#!/bin/bash
# r02aj (1 GPU): MIS rays of a scene without area lights through the any-hit kernel: image tests, C5 frame rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_host.py tests/test_refplugin.py -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 6 | tee gpurun_out/r02aj_pytest.txt
timeout 600 python tools/render_env_bench.py 32 2500 2000 0 2 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02aj_env_capi.txt
