#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu7.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu7.log | cut -c1-250
