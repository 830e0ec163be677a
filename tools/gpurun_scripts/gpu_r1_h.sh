#!/bin/bash
# distributed inverse: thread + IPC tests, full GPU suite, 2-GPU bench with sharded e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu5.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu5.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_n50k_g2_dist.log 2>&1
echo "bench g2 exit $?"; grep '^{' gpurun_out/bench_n50k_g2_dist.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], json.dumps(d['e2e']), json.dumps(d['setup_s']))"
tail -5 gpurun_out/bench_n50k_g2_dist.log | grep -v '^{' | cut -c1-400
