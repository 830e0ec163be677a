#!/bin/bash
# Round 2, call H (1 GPU): suite after the wide-N int8 MMAs / tile walk / pinv rank rule, then GEMM and one-call timings.
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -12 $O/pytest_gpu.log | cut -c1-250
for n in 4096 8192 16384; do
  timeout 200 python tools/emulated_gemm_bench.py $n 8 2>&1 | tail -1 | tee -a $O/emulated_bench.jsonl
done
timeout 300 python tools/e2e_only.py 3 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e.log
VGP_OPT_GEMM_EMULATE_MIN=1024 timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e_min1024.log
VGP_OPT_GEMM_EMULATE_MIN=4096 timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e_min4096.log
