#!/bin/bash
# 8-GPU box: driver-style bench at N=4 and N=8 (n = 50k), then cfg5 (n = 100k, N=8)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
free -g > gpurun_out/host_mem8.txt
run() { # N tag extra...
  N=$1; TAG=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 100 --warmup 3 "$@" > gpurun_out/bench_${TAG}.log 2>&1
  echo "bench $TAG exit $?"; grep '^{' gpurun_out/bench_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], json.dumps(d['e2e']), json.dumps(d['setup_s']))"
  tail -5 gpurun_out/bench_${TAG}.log | grep -v '^{' | cut -c1-400
}
run 8 n50k_g8
run 4 n50k_g4
run 8 n100k_g8 --n 100000
run 8 n50k_g8_nccl --exchange nccl --no-e2e
