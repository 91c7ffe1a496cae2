#!/bin/bash
# GPU session 9 (1 GPU): parity tests, smoke, bench (both arms), ncu launch list + full capture of the default closest-hit kernel,
# then BASELINE render configs C3/C4 (full spp) and C5 (64 of 256 spp) through the host CLI on one GPU.
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-200
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json | cut -c1-400
tail -5 gpurun_out/bench_err.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --render-spp 2 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full capture of the closest-hit kernel at the bench's ray count"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traceCoop -s 2 -c 1 -f -o gpurun_out/prof_final_16M \
    python tools/sweep2.py 16777216 5 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
echo "== render configs, 1 GPU"
timeout 600 python tools/render_configs.py 1 c3,c4 1.0 2>&1 | tail -2 | cut -c1-420
cp gpurun_out/render_configs_g1.json gpurun_out/render_configs_g1_c3c4.json
timeout 900 python tools/render_configs.py 1 c5 0.25 2>&1 | tail -1 | cut -c1-600
cp gpurun_out/render_configs_g1.json gpurun_out/render_configs_g1_c5.json
ls gpurun_out
