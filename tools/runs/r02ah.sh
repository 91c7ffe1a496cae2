#!/bin/bash
# r02ah (1 GPU): builder work arrays from the context's arena: build phases (10 M x4, 1 M host/device alternating), GPU suite, builder memcheck, bench N=1
mkdir -p gpurun_out
( timeout 300 python tools/build_cold.py 2500 2000 0,0,0,0; timeout 300 python tools/build_cold.py 1000 500 2,0,0,2,0 ) 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02ah_build_cold.txt | grep -E "build [0-9]|allocations"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 6 | tee gpurun_out/r02ah_pytest_gpu.txt
S=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $S --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_trace_gpu.py -x -q -k "device_sah_builder or device_builder_edge" > gpurun_out/r02ah_sanitizer_builder_memcheck.log 2>&1; echo "builder memcheck: exit $? ; $(grep -E 'ERROR SUMMARY' gpurun_out/r02ah_sanitizer_builder_memcheck.log | tail -n 1)" | tee gpurun_out/r02ah_sanitizer_summary.txt
tail -n 30 gpurun_out/r02ah_sanitizer_builder_memcheck.log > gpurun_out/r02ah_sanitizer_builder_memcheck.log.tail; rm -f gpurun_out/r02ah_sanitizer_builder_memcheck.log
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02ah_bench_err.txt > gpurun_out/r02ah_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ah_bench_n1.json').read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'c3', round(d['extra']['render_c3']['msamples_s'],1), 'c4', round(d['extra']['render_c4']['msamples_s'],1), 'c1', d['extra']['cornell_c1']['gpu_seconds_all_runs'], 'builders', d['extra']['builders']['device_sah']['build_s'], d['extra']['builders']['wide_bvh_bytes_equal'])
PY
