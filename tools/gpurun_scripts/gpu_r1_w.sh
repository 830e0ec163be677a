#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_expquad_dense.py -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_w.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_w.log | cut -c1-250
timeout 300 python tools/calc_h_bench.py 2>&1 | tail -3
