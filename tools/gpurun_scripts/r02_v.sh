#!/bin/bash
# Round 2, call V (4 GPUs): int8 threshold of distributed products at 4 ranks.
mkdir -p gpurun_out/r02v
O=gpurun_out/r02v
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
    tools/dist_inverse_bench.py 50000 dist_emulate_min=-1 dist_emulate_min=1024 dist_emulate_min=4096 \
    dist_emulate_min=8192 > $O/dist_inverse_bench.jsonl 2> $O/err.txt
echo "rc=$?" | tee $O/rc.txt
tail -c 300 $O/err.txt
cat $O/dist_inverse_bench.jsonl | cut -c1-400
