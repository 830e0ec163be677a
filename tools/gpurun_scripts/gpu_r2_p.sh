#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-elbo"
timeout 900 ncu --set full --clock-control none -k regex:'trigemv_kernel<false>|trigemv_kernel<0>|trigemv' -c 12 -f -o gpurun_out/prof_trigemv $CMD > gpurun_out/ncu_trigemv.log 2>&1
echo "ncu exit $?"; grep '^{' gpurun_out/ncu_trigemv.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['lazy_column']['lazy_factor']['selection_head']))"
ncu -i gpurun_out/prof_trigemv.ncu-rep --page raw --csv > gpurun_out/prof_trigemv_raw.csv 2>/dev/null
ls -la gpurun_out/prof_trigemv*
