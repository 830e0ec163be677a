#!/bin/bash
# overlapped lower-triangle H2D in the one-call lazy path: parity, then e2e; GEMM shape sweep vs cuBLAS
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_greedy.py -m gpu -q --maxfail=10 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_l.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_l.log | cut -c1-300
timeout 900 python bench.py --no-cpu --no-elbo --no-lazy > gpurun_out/bench_n50k_overlap.log 2>&1
echo "bench exit $?"; grep '^{' gpurun_out/bench_n50k_overlap.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac']); print(json.dumps(d['e2e']))"
tail -5 gpurun_out/bench_n50k_overlap.log | grep -v '^{' | cut -c1-400
timeout 900 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1
echo "sweep exit $?"; cat gpurun_out/gemm_sweep.log | cut -c1-250
