#!/bin/bash
# r02f (1 GPU): builder byte test after the canonical index split, full GPU suite, smoke, bench N=1 both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02f_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 | tee gpurun_out/r02f_smoke.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02f_bench_err.txt > gpurun_out/r02f_bench_n1.json ) 2>&1 | tail -n 4
cut -c1-400 gpurun_out/r02f_bench_n1.json; tail -n 5 gpurun_out/r02f_bench_err.txt
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/r02f_bench_err.txt > gpurun_out/r02f_bench_reference_arm.json ) 2>&1 | tail -n 4
cut -c1-300 gpurun_out/r02f_bench_reference_arm.json
