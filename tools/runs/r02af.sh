#!/bin/bash
# r02af (1 GPU): scratch pool + tree re-use: rebuild phases, the GPU suite, memcheck on the builder / a render / the new film sums,
# the deep-stack kernel, C1 through the host repeatedly
mkdir -p gpurun_out
timeout 300 python tools/build_cold.py 2500 2000 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02af_build_cold.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee gpurun_out/r02af_pytest_gpu.txt
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool command...
  name=$1; tool=$2; shift 2
  timeout 600 $S --tool $tool --error-exitcode 77 --print-limit 20 "$@" > gpurun_out/r02af_sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool: exit $? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02af_sanitizer_${name}_${tool}.log | tail -n 1)"
}
( run builder memcheck python -m pytest tests/test_trace_gpu.py -x -q -k "device_sah_builder or device_builder_edge"
  run f64_and_deep memcheck python -m pytest tests/test_trace_gpu.py -x -q -k "transformed_mesh and 5-0.0 or deep_imported"
  run f64_and_deep racecheck python -m pytest tests/test_trace_gpu.py -x -q -k "transformed_mesh and 5-0.0 or deep_imported"
  run film_sums memcheck python -m pytest tests/test_render_gpu.py -x -q -k "peer_memory or cuda_ipc"
  run render_glossy memcheck python tools/render_once.py glossy 4 64 64 8 ) 2>&1 | tee gpurun_out/r02af_sanitizer_summary.txt
for f in gpurun_out/r02af_sanitizer_*.log; do tail -n 40 $f > $f.tail; rm -f $f; done
timeout 600 python bench.py --steps 3 --warmup 3 --c3-spp 64 --c4-spp 64 2> gpurun_out/r02af_bench_err.txt > gpurun_out/r02af_bench_short.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02af_bench_short.json').read())
print('value', round(d['value'],1), 'c1', d['extra']['cornell_c1']['gpu_seconds_all_runs'], 'builders', d['extra']['builders']['device_sah']['build_s'], d['extra']['builders']['wide_bvh_bytes_equal'])
PY
