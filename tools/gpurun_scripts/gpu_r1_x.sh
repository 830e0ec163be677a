#!/bin/bash
# 2 GPUs: multi-process tests (IPC dist inverse + ShardedPlacer, peer exchange), then the driver-style bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist_inverse.py tests/test_gpu_greedy.py -m gpu -q --maxfail=10 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_x.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_x.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_n50k_g2_latest.log 2>&1
echo "bench g2 exit $?"; grep '^{' gpurun_out/bench_n50k_g2_latest.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], json.dumps(d['e2e']), json.dumps(d['setup_s']))"
tail -4 gpurun_out/bench_n50k_g2_latest.log | grep -v '^{' | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-300
