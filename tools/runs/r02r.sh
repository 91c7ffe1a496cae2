#!/bin/bash
# r02r (1 GPU): final build: full GPU suite, smoke, sanitizer pass, ncu of the final traversal kernel (C2) for traffic.json, bench N=1 both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee gpurun_out/r02r_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02r_smoke.txt
bash tools/runs/r02_sanitize.sh
cp gpurun_out/r02_sanitizer_summary.txt gpurun_out/r02r_sanitizer_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traceCoopPair -s 2 -c 1 -o gpurun_out/r02r_trace_c2 python tools/sweep4.py 16777216 5 > gpurun_out/r02r_ncu_trace_c2.log 2>&1
bash tools/ncu_to_csv.sh gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02r_bench_err.txt > gpurun_out/r02r_bench_n1.json; wc -l gpurun_out/r02r_bench_n1.json; cut -c1-250 gpurun_out/r02r_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/r02r_bench_err.txt > gpurun_out/r02r_bench_reference_arm.json; cut -c1-200 gpurun_out/r02r_bench_reference_arm.json
