#!/bin/bash
# Round 2, call J (1 GPU): lowrank tests again, e2e wall after the no-cudaFree change, ncu launch list of the ELBO step.
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
timeout 600 python -m pytest tests/test_gpu_lowrank.py tests/test_gpu_lazy.py tests/test_gpu_elbo.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -5 $O/pytest.log | cut -c1-250
timeout 300 python tools/e2e_only.py 4 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e.log
timeout 300 python tools/elbo_steps.py --steps 3 > $O/elbo_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/elbo_launches.csv \
    python tools/elbo_steps.py --steps 3 > $O/elbo_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/rc.txt
tail -3 $O/elbo_plain.log
