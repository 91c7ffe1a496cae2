#!/bin/bash
# r02n (2 GPUs): CLI phases after geometry sharing / overlapped NCCL init / parallel teardown; host + refplugin GPU tests; cost of CUDA start-up itself
mkdir -p gpurun_out /tmp/rc
python -c "
import time, ctypes
t0=time.perf_counter(); rt=ctypes.CDLL('libcudart.so.12'); rt.cudaFree(None); print('first cudaFree(0) in a fresh process: %.3f s' % (time.perf_counter()-t0))
" 2>&1 | tee gpurun_out/r02n_cuda_startup.txt
timeout 600 python -m pytest tests/test_host.py tests/test_refplugin.py -x -q -m gpu 2>&1 | grep -v "^\[INFO\]" | tail -n 6 | tee gpurun_out/r02n_pytest_hosts.txt
python -c "
from spica_b200 import scenes
scenes.write_cornell('/tmp/rc', 1920, 1080, 256, 16, variant='diffuse', name='c3')
scenes.write_envscene('/tmp/rc', 3840, 2160, 16, 16, name='c5', nu=2500, nv=2000)
" 2>&1 | tail -n 3
export SPICA_TIMING=1
cd spica_b200/bin
for c in c3 c5; do for g in 1 2; do
  echo "== $c gpus $g"
  ./spica -i /tmp/rc/$c.xml -o /tmp/rc/${c}_out --gpus $g --seed 1 2>&1 | grep -E "TIME|rendered|BVH|rror" 
done; done 2>&1 | tee ../../gpurun_out/r02n_cli_phases.txt
cd ../..
timeout 600 python tools/render_configs.py 2 c3,c4 0.25 2>&1 | tail -n 2 | cut -c1-330 | tee gpurun_out/r02n_render_configs_g2.txt
timeout 600 python tools/render_configs.py 1 c3,c4 0.25 2>&1 | tail -n 2 | cut -c1-330 | tee gpurun_out/r02n_render_configs_g1.txt
