#!/bin/bash
# pair-CTA GEMM as default (+ in-place guard), new kernel-matrix builder: full GPU suite, kernel bench, e2e x3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu6.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu6.log | cut -c1-300
timeout 600 python tools/kernel_bench.py 50000 > gpurun_out/kernel_bench.log 2>&1; echo "kernel bench exit $?"; cat gpurun_out/kernel_bench.log | cut -c1-220
timeout 600 python tools/kernel_bench.py 10768 >> gpurun_out/kernel_bench.log 2>&1; tail -12 gpurun_out/kernel_bench.log | grep expquad | cut -c1-220
timeout 600 python tools/e2e_only.py 4 auto 2>&1 | grep overlap
