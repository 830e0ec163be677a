#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'kernel_rect|kernel_sym' -c 3 -f -o gpurun_out/prof_kernel_matrix python tools/kernel_bench.py 50000 > gpurun_out/ncu_kernel_matrix.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_kernel_matrix.log
ncu -i gpurun_out/prof_kernel_matrix.ncu-rep --page raw --csv > gpurun_out/prof_kernel_matrix_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_kernel_matrix.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_kernel_matrix_sass.csv 2>/dev/null
ls -la gpurun_out/prof_kernel_matrix*
