#!/bin/bash
# smoke(), then the default bench line and the reference arm exactly as the driver runs them, timed
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
( time python bench.py --impl reference ) > gpurun_out/bench_ref_default.log 2>&1; echo "ref exit $?"; grep real gpurun_out/bench_ref_default.log
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; grep real gpurun_out/bench_default.log
grep '^{' gpurun_out/bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks']); print(json.dumps(d['e2e'])); print(json.dumps(d['setup_s'])); print(json.dumps(d.get('elbo'))[:400]); print(json.dumps(d.get('lazy_column'))[:600])"
