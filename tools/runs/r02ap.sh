#!/bin/bash
# r02ap (1 GPU): the host after the parallel mesh loading / flatten: its GPU tests (incl. the textured scene: normals + uvs), and the phases of a C5 job at 8 spp
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_host.py -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 3 | tee gpurun_out/r02aq_pytest_host.txt
python -c "
from spica_b200 import scenes
scenes.write_envscene('/tmp/rc', 3840, 2160, 8, 16, name='c5', nu=2500, nv=2000)
" 2>&1 | tail -n 1
( cd spica_b200/bin; SPICA_TIMING=1 timeout 40 ./spica -i /tmp/rc/c5.xml -o /tmp/rc/c5_out --gpus 1 --seed 1 2>&1 | grep -E "TIME|rendered|rror|BVH" ) | tee gpurun_out/r02aq_cli_phases_c5.txt
