#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=5 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu8.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu8.log | cut -c1-200
timeout 600 python tools/e2e_only.py 4 auto 2>&1 | grep overlap
