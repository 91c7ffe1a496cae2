#!/bin/bash
# GPU session 10 (1 GPU): final verification of HEAD: parity tests, smoke, bench (both arms), launch list, full ncu capture at the bench's configuration
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-200
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json | cut -c1-300
tail -5 gpurun_out/bench_err.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --render-spp 2 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-120
echo "== ncu full capture of the closest-hit kernel at the bench's ray count and tree"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traceCoop -s 2 -c 1 -f -o gpurun_out/prof_final_16M \
    python tools/sweep2.py 16777216 5 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
