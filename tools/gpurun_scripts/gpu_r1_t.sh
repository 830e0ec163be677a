#!/bin/bash
# full default bench line (N = 1), then the ncu launch list of the same command shape
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_n50k_r1final.log 2>&1
echo "bench exit $?"; grep '^{' gpurun_out/bench_n50k_r1final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac']); print(json.dumps(d['e2e'])); print(json.dumps(d['setup_s'])); print(json.dumps(d.get('elbo'))[:600]); print(json.dumps(d.get('cpu_baseline'))[:300])"
tail -3 gpurun_out/bench_n50k_r1final.log | grep -v '^{' | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_r1final.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_r1final.log | cut -c1-400
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-elbo --no-lazy"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_n50k_b.csv $CMD > gpurun_out/ncu50k_list_b.log 2>&1
echo "launch list exit $?"
