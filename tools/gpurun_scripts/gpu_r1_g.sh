#!/bin/bash
# peer-memory exchange: single-device multi-shard tests, 2-process IPC test, 2-GPU bench (peer vs nccl)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_greedy.py -m gpu -q -k "peer" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -15 gpurun_out/pytest_peer.log | cut -c1-300
for ex in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 --exchange $ex > gpurun_out/bench_n50k_g2_$ex.log 2>&1
echo "bench g2 $ex exit $?"; grep '^{' gpurun_out/bench_n50k_g2_$ex.log | cut -c1-160; tail -3 gpurun_out/bench_n50k_g2_$ex.log | grep -v '^{' | cut -c1-300
done
timeout 900 python bench.py --n 10768 --steps 50 --warmup 3 --no-cpu > gpurun_out/bench4_n10k.log 2>&1; echo "bench10k exit $?"
grep '^{' gpurun_out/bench4_n10k.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d.get('elbo'))[:600])"
