#!/bin/bash
# r02ad (1 GPU): the final build: full GPU suite (new: peer / IPC film sums on one GPU, deep stack, double-precision scenes), smoke,
# bench N=1 + reference arm, and the ncu launch list of a short bench.py run
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee gpurun_out/r02ad_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02ad_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02ad_bench_err.txt > gpurun_out/r02ad_bench_n1.json; wc -l gpurun_out/r02ad_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r02ad_ref_err.txt > gpurun_out/r02ad_bench_reference_arm.json; cut -c1-300 gpurun_out/r02ad_bench_reference_arm.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ad_bench_n1.json').read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'instr', round(d['roofline']['instruction']['frac'],3))
for k in ('render_c3','render_c4'):
    r=d['extra'][k]; print(' ', k, round(r['msamples_s'],1), 'roof frac', round(r['roofline']['frac'],3))
print(' c1', d['extra']['cornell_c1']['gpu_seconds_all_runs'], 'builders', d['extra']['builders']['device_sah']['build_s'], d['extra']['builders']['wide_bvh_bytes_equal'])
print(' parity', d['parity']['closest_incoherent']['prim_id_mismatches'], d['parity']['closest_primary']['prim_id_mismatches'], d['parity']['any_hit']['flag_mismatches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02ad_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --c3-spp 16 --c4-spp 16 > gpurun_out/r02ad_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02ad_launches_bench.csv 2>&1 | head -n 30 | tee gpurun_out/r02ad_launch_share_bench.txt
