#!/bin/bash
# final 8-GPU lines: n = 50k (driver style), cfg5 n = 100k, and the distributed-inverse thresholds experiment
mkdir -p gpurun_out
run() { TAG=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 100 --warmup 3 "$@" > gpurun_out/bench_${TAG}.log 2>&1
  echo "bench $TAG exit $?"; grep '^{' gpurun_out/bench_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline'].get('traffic')); print(json.dumps(d['e2e'])[:700]); print(json.dumps(d['setup_s'])); print(json.dumps(d.get('elbo'))[:300])"
  tail -3 gpurun_out/bench_${TAG}.log | grep -v '^{' | cut -c1-300
}
run n50k_g8_final
VGP_BENCH_N=100000 run n100k_g8_final
VGP_DIST_MIN_TILES=96 VGP_DIST_MIN_K=256 run n50k_g8_thresholds --no-elbo
