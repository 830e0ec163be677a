#!/bin/bash
# ELBO step breakdown + ncu capture of the downdate kernel at the bench workload (n = 50 000)
mkdir -p gpurun_out
timeout 600 python tools/elbo_profile.py 8 > gpurun_out/elbo_profile.log 2>&1; echo "elbo profile exit $?"; cat gpurun_out/elbo_profile.log
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-elbo"
timeout 600 $CMD > gpurun_out/ncu50k_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:downdate -s 3 -c 2 -f -o gpurun_out/prof_downdate_n50k $CMD > gpurun_out/ncu50k_full.log 2>&1
echo "downdate n50k capture exit $?"
ncu -i gpurun_out/prof_downdate_n50k.ncu-rep --page raw --csv > gpurun_out/prof_downdate_n50k_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_n50k.csv $CMD > gpurun_out/ncu50k_list.log 2>&1
echo "launch list exit $?"
