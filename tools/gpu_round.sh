#!/bin/bash
# One GPU-box session: parity tests, bench, sweep, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
echo "== sweep"; timeout 600 python tools/sweep.py 4194304 2>&1 | tee gpurun_out/sweep.txt | tail -30
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --rays 4194304 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePersistent -s 1 -c 1 -f -o gpurun_out/prof \
    python bench.py --rays 4194304 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
