#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_expquad_dense.py tests/test_gpu_gp.py tests/test_gpu_elbo.py tests/test_gpu_dist_inverse.py tests/test_gpu_lazy.py -m gpu -q --maxfail=20 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_s.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_s.log | cut -c1-300
timeout 600 python tools/elbo_profile.py 6 > gpurun_out/elbo_profile3.log 2>&1; echo "elbo profile exit $?"; cat gpurun_out/elbo_profile3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_elbo3.csv python tools/elbo_profile.py 3 > gpurun_out/ncu_elbo3.log 2>&1
echo "launch list exit $?"
timeout 600 python tools/e2e_only.py 3 auto 2>&1 | grep overlap
