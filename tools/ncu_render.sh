#!/bin/bash
# ncu launch list (durations) of the render loop alone: Cornell diffuse (c3) and glossy (c4) at 1920x1080, 4 spp, via the host CLI
mkdir -p gpurun_out /tmp/rc
python - <<'PY'
from spica_b200 import scenes
scenes.write_cornell("/tmp/rc", 1920, 1080, 8, 16, variant="diffuse", name="c3")
scenes.write_cornell("/tmp/rc", 1920, 1080, 8, 16, variant="glossy", name="c4")
PY
for c in c3 c4; do
  (cd spica_b200/bin && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file ../../gpurun_out/launches_render_$c.csv \
      ./spica -i /tmp/rc/$c.xml -o /tmp/rc/${c}_out --seed 1 > ../../gpurun_out/ncu_render_$c.log 2>&1)
  tail -3 gpurun_out/ncu_render_$c.log
done
