#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/elbo_profile.py 6 > gpurun_out/elbo_profile2.log 2>&1; echo "elbo profile exit $?"; cat gpurun_out/elbo_profile2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_elbo2.csv python tools/elbo_profile.py 3 > gpurun_out/ncu_elbo2.log 2>&1
echo "launch list exit $?"
