#!/bin/bash
# ncu pass (one GPU): launch list of a short bench run + full captures of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --n 10768 --steps 5 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_n10k.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:downdate -s 3 -c 2 -f -o gpurun_out/prof_downdate $CMD > gpurun_out/ncu_full_downdate.log 2>&1
echo "downdate capture exit $?"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 600 -c 2 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
$CMD > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:expquad -c 1 -f -o gpurun_out/prof_expquad $CMD > gpurun_out/ncu_full_expquad.log 2>&1
echo "expquad capture exit $?"
ls -la gpurun_out
