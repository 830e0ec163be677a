#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log; tail -12 gpurun_out/pytest_gpu3.log | cut -c1-200
timeout 600 python bench.py --n 10768 --steps 50 --warmup 3 --no-elbo --no-cpu > gpurun_out/bench3_n10k.log 2>&1; echo "bench10k exit $?"
CMD="python tools/elbo_steps.py --steps 3"
$CMD > gpurun_out/elbo_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_elbo.csv $CMD > gpurun_out/ncu_elbo.log 2>&1
echo "elbo ncu exit $?"; cat gpurun_out/elbo_plain.log
