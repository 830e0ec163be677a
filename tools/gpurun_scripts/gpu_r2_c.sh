#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elbo.py -m gpu -q --maxfail=5 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_elbo_shard.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_elbo_shard.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 3 --no-e2e > gpurun_out/bench_g2_elbo.log 2>&1
echo "bench g2 exit $?"; grep '^{' gpurun_out/bench_g2_elbo.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline'].get('traffic')); print(json.dumps(d.get('elbo')))"
tail -3 gpurun_out/bench_g2_elbo.log | grep -v '^{' | cut -c1-300
