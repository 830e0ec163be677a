#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_expquad_dense.py tests/test_gpu_gp.py tests/test_gpu_elbo.py -m gpu -q --maxfail=20 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_q.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_q.log | cut -c1-300
timeout 600 python tools/kernel_bench.py 50000 > gpurun_out/kernel_bench.log 2>&1; echo "kernel bench exit $?"; cat gpurun_out/kernel_bench.log | cut -c1-220
ncu --set full --clock-control none --import-source on -k regex:'kernel_rect|kernel_sym' -c 2 -f -o gpurun_out/prof_kernel_matrix2 python tools/kernel_bench.py 50000 > gpurun_out/ncu_kernel_matrix2.log 2>&1
ncu -i gpurun_out/prof_kernel_matrix2.ncu-rep --page raw --csv > gpurun_out/prof_kernel_matrix2_raw.csv 2>/dev/null
