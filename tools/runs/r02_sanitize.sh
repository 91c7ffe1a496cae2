#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the hot path: every traversal variant on the golden torus, the device
# BVH builder, and one small Cornell frame (diffuse: queue-order shading; glossy: classify + per-bucket shading)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool command...
  name=$1; tool=$2; shift 2
  timeout 900 $S --tool $tool --error-exitcode 77 --print-limit 20 "$@" > gpurun_out/r02_sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool: exit $? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_${name}_${tool}.log | tail -n 1)"
}
for tool in memcheck racecheck synccheck; do
  run trace $tool python -m pytest tests/test_trace_gpu.py -x -q -k "golden_torus or edge_cases or golden_cube"
  run builder $tool python -m pytest tests/test_trace_gpu.py -x -q -k "device_sah_builder or device_builder_edge"
  run render_diffuse $tool python tools/render_once.py diffuse 4 64 64 8
  run render_glossy $tool python tools/render_once.py glossy 4 64 64 8
done 2>&1 | tee gpurun_out/r02_sanitizer_summary.txt
# keep the logs small
for f in gpurun_out/r02_sanitizer_*.log; do tail -n 60 $f > $f.tail; rm -f $f; done
