#!/bin/bash
# Round 2, call A: first hardware contact of the never-run paths (int8 tcgen05 GEMM, slab solves, wide leaves).
# Tests gate the timings: a failing test skips the timings of that path.  Everything bounded by timeout.
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
VGP_TEST_EMULATED=1 timeout 420 python -m pytest tests/test_gpu_emulated_gemm.py -q --timeout 120 -p no:cacheprovider \
    > $O/emulated_tests.log 2>&1
echo "emulated tests rc=$?" | tee -a $O/rc.txt
tail -40 $O/emulated_tests.log
if grep -q " passed" $O/emulated_tests.log && ! grep -q " failed" $O/emulated_tests.log; then
  for v in 1 2; do
    for n in 4096 8192; do
        VGP_GEMM_EMULATE_VARIANT=$v timeout 120 python tools/emulated_gemm_bench.py $n 8 2>&1 | tail -1 \
            | tee -a $O/emulated_bench.jsonl
    done
  done
  for v in 1 2; do
    VGP_GEMM_EMULATE=8 VGP_GEMM_EMULATE_VARIANT=$v timeout 200 python tools/e2e_only.py 2 auto 2>&1 | grep overlap \
        | tee $O/e2e_emulated_v$v.log
  done
fi
timeout 200 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee $O/e2e_fp64_pipe.log
VGP_TEST_SLAB=1 timeout 420 python -m pytest tests/test_gpu_slab_solves.py -q --timeout 120 -p no:cacheprovider \
    > $O/slab_tests.log 2>&1
echo "slab tests rc=$?" | tee -a $O/rc.txt
tail -25 $O/slab_tests.log
if grep -q " passed" $O/slab_tests.log && ! grep -q " failed" $O/slab_tests.log; then
  for w in 512 1024; do
    VGP_TRSM_SLAB=$w timeout 200 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee $O/e2e_slab_$w.log
  done
  for w in 256 512; do
    VGP_TRSM_LEAF=$w timeout 200 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee $O/e2e_leaf_$w.log
  done
fi
