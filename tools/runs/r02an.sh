#!/bin/bash
# r02an (1 GPU): full GPU suite on the final build (after making the reference-integrator-on-GPU-accelerator test robust), twice
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^\[INFO\]" | tail -n 15 | tee gpurun_out/r02an_pytest_gpu.txt
timeout 600 python -m pytest tests/test_refplugin.py -m gpu -q -k "reference_integrator_on_the_gpu_accelerator" --durations=3 2>&1 | grep -v "^\[INFO\]" | tail -n 8 | tee -a gpurun_out/r02an_pytest_gpu.txt
