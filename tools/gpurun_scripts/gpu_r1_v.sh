#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_cov_producer.py -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_v.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_v.log | cut -c1-250
