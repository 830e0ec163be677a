#!/bin/bash
# end-of-round check on one GPU: smoke, full GPU suite, default bench line and reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_final.log | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_final.log 2>&1; echo "ref exit $?"
timeout 1500 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['gpu_launches']); print(json.dumps(d['e2e'])[:500]); print(json.dumps(d['setup_s'])); e=d.get('elbo'); print(e['value'], e['ms_per_step'], e['tflops']); print(json.dumps(d['cpu_baseline'])[:200])"
