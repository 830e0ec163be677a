#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 120 -p no:cacheprovider > gpurun_out/pytest_tma2.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_tma2.log | cut -c1-300
timeout 600 python tools/gemm_sweep.py 2>&1 | grep -o "'m': [0-9]*, 'n': [0-9]*, 'k': [0-9]*\|'ours_tflops': [0-9.]*\|'cublas_tflops': [0-9.]*" | paste - - - | head -20
cp gpurun_out/gemm_sweep.json gpurun_out/gemm_sweep_tma.json
timeout 600 python tools/e2e_only.py 3 auto 2>&1 | grep overlap
timeout 600 python tools/elbo_profile.py 5 2>&1 | tail -3
