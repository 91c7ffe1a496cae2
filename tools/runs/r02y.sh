#!/bin/bash
# r02y (1 GPU): double-precision-vertex scenes on the pooled kernels (rounded pre-test): tests, A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_gpu.py -m gpu -x -q 2>&1 | grep -v "^\[INFO\]" | tail -n 6 | tee gpurun_out/r02y_pytest_trace.txt
AB=$PWD/spica_b200/lib_ab/libspica_b200.so
for lib in new old; do
  for tf in 0 1; do
    echo "== lib=$lib transform=$tf"
    if [ $lib = old ]; then export SPICA_B200_LIB=$AB; else unset SPICA_B200_LIB; fi
    if [ $tf = 1 ]; then export SPB_TRANSFORM=1; else unset SPB_TRANSFORM; fi
    timeout 600 python tools/sweep4.py 16777216 5,2 2>&1 | grep -v "^\[INFO\]"
  done
done | tee gpurun_out/r02y_sweep_f64.txt
unset SPICA_B200_LIB SPB_TRANSFORM
timeout 600 python tools/render_env_bench.py 32 2500 2000 0 2 2>&1 | grep -v "^\[INFO\]" | tee gpurun_out/r02y_env_capi.txt
