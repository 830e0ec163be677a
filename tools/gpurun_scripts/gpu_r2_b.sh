#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_one.py 8192 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_tma' -c 2 -f -o gpurun_out/prof_gemm_tma_nt python tools/gemm_one.py 8192 1 > gpurun_out/ncu_gemm_tma_nt.log 2>&1; echo "ncu nt exit $?"
ncu --set full --clock-control none --import-source on -k regex:'gemm_tma' -c 2 -f -o gpurun_out/prof_gemm_tma_nn python tools/gemm_one.py 8192 0 > gpurun_out/ncu_gemm_tma_nn.log 2>&1; echo "ncu nn exit $?"
for v in nt nn; do ncu -i gpurun_out/prof_gemm_tma_$v.ncu-rep --page raw --csv > gpurun_out/prof_gemm_tma_${v}_raw.csv 2>/dev/null; done
ls -la gpurun_out/prof_gemm_tma*
