#!/bin/bash
# lazy-column formulation: parity tests, then the full bench line (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_greedy.py -m gpu -q --maxfail=10 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_lazy.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_lazy.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/bench_n50k_lazy.log 2>&1
echo "bench exit $?"; grep '^{' gpurun_out/bench_n50k_lazy.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac']); print(json.dumps(d['e2e'])); print(json.dumps(d.get('lazy_column'))); print(json.dumps(d.get('elbo'))[:700])"
tail -5 gpurun_out/bench_n50k_lazy.log | grep -v '^{' | cut -c1-400
