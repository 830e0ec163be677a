#!/bin/bash
# first GPU pass: parity tests, FP64 peak, small + full bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
free -g > gpurun_out/host_mem.txt 2>&1; nproc >> gpurun_out/host_mem.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/measure_fp64_peak.py > gpurun_out/fp64.log 2>&1; echo "fp64 exit $?"; tail -2 gpurun_out/fp64.log
timeout 600 python bench.py --n 10768 --steps 50 --warmup 3 > gpurun_out/bench_n10k.log 2>&1; echo "bench10k exit $?"; tail -c 1500 gpurun_out/bench_n10k.log
timeout 1200 python bench.py > gpurun_out/bench_n50k.log 2>&1; echo "bench50k exit $?"; tail -c 2500 gpurun_out/bench_n50k.log
