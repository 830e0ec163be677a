#!/bin/bash
# 8-GPU box: cfg5 (n = 100k, k = 100, N = 8)
mkdir -p gpurun_out
export VGP_BENCH_N=100000
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/bench_n100k_g8.log 2>&1
echo "bench exit $?"; grep '^{' gpurun_out/bench_n100k_g8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], json.dumps(d['e2e']), json.dumps(d['setup_s']))"
tail -5 gpurun_out/bench_n100k_g8.log | grep -v '^{' | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv | head -3
