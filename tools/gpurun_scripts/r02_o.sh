#!/bin/bash
# Round 2, call O (1 GPU): whole suite after the coalesced int8 epilogue, pageable-copy helper, runtime tests; timings.
mkdir -p gpurun_out/r02o
O=gpurun_out/r02o
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -8 $O/pytest_gpu.log | cut -c1-250
for n in 4096 8192; do timeout 200 python tools/emulated_gemm_bench.py $n 8 2>&1 | tail -1 | tee -a $O/emu.log; done
timeout 300 python tools/e2e_only.py 3 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e.log
timeout 300 python tools/elbo_steps.py --steps 6 2>&1 | tail -2 | tee $O/elbo.log
