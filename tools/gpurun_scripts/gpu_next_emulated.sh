#!/bin/bash
# First GPU call for the experimental int8-tensor-core GEMM (csrc/emulated.cu): numerics, then speed, then the one-call
# placement with the factorisation's large products routed through it.  Every step is bounded (the kernels trap after
# 4 s on a lost barrier instead of hanging).
mkdir -p gpurun_out
VGP_TEST_EMULATED=1 timeout 600 python -m pytest tests/test_gpu_emulated_gemm.py -x -q --timeout 200 -p no:cacheprovider \
    > gpurun_out/emulated_tests.log 2>&1
tail -15 gpurun_out/emulated_tests.log
for v in 1 2; do
    for n in 4096 8192; do
        VGP_GEMM_EMULATE_VARIANT=$v timeout 200 python tools/emulated_gemm_bench.py $n 8 2>&1 | tail -1 \
            | tee -a gpurun_out/emulated_bench.jsonl
    done
done
timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee gpurun_out/e2e_fp64_pipe.log
for v in 1 2; do
    VGP_GEMM_EMULATE=8 VGP_GEMM_EMULATE_VARIANT=$v timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep overlap \
        | tee gpurun_out/e2e_emulated_v$v.log
done
# slab kernels for the triangular solves (csrc/slab.cu)
VGP_TEST_SLAB=1 timeout 600 python -m pytest tests/test_gpu_slab_solves.py -x -q --timeout 240 -p no:cacheprovider \
    > gpurun_out/slab_tests.log 2>&1
tail -15 gpurun_out/slab_tests.log
for w in 512 1024 2048; do
    VGP_TRSM_SLAB=$w timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee gpurun_out/e2e_slab_$w.log
done
# wide leaves for the triangular solves (dense.cu, VGP_TRSM_LEAF)
for w in 256 512; do
    VGP_TRSM_LEAF=$w timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | tee gpurun_out/e2e_leaf_$w.log
done
