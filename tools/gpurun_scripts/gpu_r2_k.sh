#!/bin/bash
# final N-GPU line, driver style (N from the first argument)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29580+N)) bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n50k_g${N}_final.log 2>&1
echo "bench g$N exit $?"; grep '^{' gpurun_out/bench_n50k_g${N}_final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline'].get('traffic')); print(json.dumps(d['e2e']['value']), json.dumps(d['e2e']['seconds'])); print(json.dumps(d['setup_s'])); print(json.dumps(d.get('elbo'))[:200])"
tail -3 gpurun_out/bench_n50k_g${N}_final.log | grep -v '^{' | cut -c1-300
