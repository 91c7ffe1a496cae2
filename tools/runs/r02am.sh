#!/bin/bash
# r02am (1 GPU): last check of the final build: full GPU suite, smoke, bench N=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee gpurun_out/r02am_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02am_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02am_bench_err.txt > gpurun_out/r02am_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02am_bench_n1.json').read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'c3', round(d['extra']['render_c3']['msamples_s'],1), 'c4', round(d['extra']['render_c4']['msamples_s'],1), 'c1', d['extra']['cornell_c1']['gpu_seconds_all_runs'], 'builders', d['extra']['builders']['device_sah']['build_s'], d['extra']['builders']['wide_bvh_bytes_equal'])
print(' parity', d['parity']['closest_incoherent']['prim_id_mismatches'], d['parity']['closest_primary']['prim_id_mismatches'], d['parity']['any_hit']['flag_mismatches'])
PY
