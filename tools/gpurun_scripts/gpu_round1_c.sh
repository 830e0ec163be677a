#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log; tail -30 gpurun_out/pytest_gpu2.log
timeout 1500 python bench.py > gpurun_out/bench2_n50k.log 2>&1; echo "bench exit $?"; tail -c 3000 gpurun_out/bench2_n50k.log
