#!/bin/bash
# r02q (8 GPUs): phases of the CLI on 8 GPUs (C3, 1024 spp), with NCCL's defaults and with a lean bring-up (no NVLS, 8 channels)
mkdir -p gpurun_out /tmp/rc
python -c "
from spica_b200 import scenes
scenes.write_cornell('/tmp/rc', 1920, 1080, 1024, 16, variant='diffuse', name='c3')
" 2>&1 | tail -n 2
export SPICA_TIMING=1
cd spica_b200/bin
for mode in default lean default lean; do
  echo "== c3 gpus 8 ($mode)"
  if [ $mode = lean ]; then export NCCL_NVLS_ENABLE=0 NCCL_MAX_NCHANNELS=8; else unset NCCL_NVLS_ENABLE NCCL_MAX_NCHANNELS; fi
  ./spica -i /tmp/rc/c3.xml -o /tmp/rc/c3_out --gpus 8 --seed 1 2>&1 | grep -E "TIME|rendered|rror|per-GPU"
done 2>&1 | tee ../../gpurun_out/r02q_cli_phases_g8.txt
echo "== c3 gpus 1"; ./spica -i /tmp/rc/c3.xml -o /tmp/rc/c3_out --gpus 1 --seed 1 2>&1 | grep -E "TIME|rendered|rror" | tee -a ../../gpurun_out/r02q_cli_phases_g8.txt
