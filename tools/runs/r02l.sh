#!/bin/bash
# r02l (2 GPUs): where does the CLI's wall time go?  (SPICA_TIMING=1 phase laps; 1 and 2 GPUs, C3 at 256 spp; C5-size scene at 16 spp)
mkdir -p gpurun_out /tmp/rc
python -c "
from spica_b200 import scenes
scenes.write_cornell('/tmp/rc', 1920, 1080, 256, 16, variant='diffuse', name='c3')
scenes.write_envscene('/tmp/rc', 3840, 2160, 16, 16, name='c5', nu=2500, nv=2000)
" 2>&1 | tail -n 3
ls -la /tmp/rc | head
export SPICA_TIMING=1
cd spica_b200/bin
for c in c3 c5; do for g in 1 2; do
  echo "== $c gpus $g"
  s=$(date +%s.%N)
  ./spica -i /tmp/rc/$c.xml -o /tmp/rc/${c}_out --gpus $g --seed 1 2>&1 | grep -E "TIME|rendered|BVH|rror" 
  e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc) s"
done; done 2>&1 | tee ../../gpurun_out/r02l_cli_phases.txt
