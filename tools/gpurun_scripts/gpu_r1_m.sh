#!/bin/bash
mkdir -p gpurun_out
python tools/e2e_only.py 2 auto > gpurun_out/e2e_overlap.log 2>&1; echo "e2e exit $?"
VGP_H2D_OVERLAP=0 python tools/e2e_only.py 2 auto >> gpurun_out/e2e_overlap.log 2>&1; echo "e2e exit $?"
grep overlap gpurun_out/e2e_overlap.log
python tools/gemm_one.py 4096 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm|dgemm|cutlass|cublas|sm90|sm100|sm80' -c 8 -f -o gpurun_out/prof_gemm_vs_cublas python tools/gemm_one.py 4096 1 > gpurun_out/ncu_gemm_vs.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_gemm_vs.log
ncu -i gpurun_out/prof_gemm_vs_cublas.ncu-rep --page raw --csv > gpurun_out/prof_gemm_vs_cublas_raw.csv 2>/dev/null
ls -la gpurun_out/prof_gemm_vs_cublas*
