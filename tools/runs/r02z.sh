#!/bin/bash
# r02z (2 GPUs): film sum over peer memory (spb_film_reduce_peers) in the one-process hosts: tests, CLI phases peers vs nccl
mkdir -p gpurun_out /tmp/rc
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_host.py tests/test_refplugin.py -x -q -m gpu 2>&1 | grep -v "^\[INFO\]" | tail -n 12 | tee gpurun_out/r02z_pytest.txt
python -c "
from spica_b200 import scenes
scenes.write_cornell('/tmp/rc', 1920, 1080, 1024, 16, variant='diffuse', name='c3')
" 2>&1 | tail -n 2
export SPICA_TIMING=1
( cd spica_b200/bin
for mode in peers nccl peers nccl; do
  echo "== c3 gpus 2 ($mode)"
  export SPICA_FILM_REDUCE=$mode
  ./spica -i /tmp/rc/c3.xml -o /tmp/rc/c3_out --gpus 2 --seed 1 2>&1 | grep -E "TIME|rendered|rror|per-GPU"
done
unset SPICA_FILM_REDUCE
echo "== c3 gpus 1"; ./spica -i /tmp/rc/c3.xml -o /tmp/rc/c3_out --gpus 1 --seed 1 2>&1 | grep -E "TIME|rendered|rror" ) 2>&1 | tee gpurun_out/r02z_cli_phases_g2.txt
