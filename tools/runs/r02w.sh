#!/bin/bash
# r02w (1 GPU): final build again: full GPU suite in two orders, smoke, bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 5 | tee gpurun_out/r02w_pytest_gpu.txt
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_host.py tests/test_refplugin.py -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 3 | tee -a gpurun_out/r02w_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02w_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02w_bench_err.txt > gpurun_out/r02w_bench_n1.json; wc -l gpurun_out/r02w_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench_n1.json').read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'instr', round(d['roofline']['instruction']['frac'],3))
for k in ('render_c3','render_c4'):
    r=d['extra'][k]; print(' ', k, round(r['msamples_s'],1), 'roof frac', round(r['roofline']['frac'],3), 'traffic', r['roofline']['traffic'])
print(' c1', d['extra']['cornell_c1']['gpu_seconds_all_runs'], 'builders', d['extra']['builders']['device_sah']['build_s'], d['extra']['builders']['wide_bvh_bytes_equal'])
print(' parity', d['parity']['closest_incoherent'], d['parity']['closest_primary']['prim_id_mismatches'], d['parity']['any_hit']['flag_mismatches'])
PY
