#!/bin/bash
# Round 2, call U (8 GPUs): threshold of the int8 path for DISTRIBUTED products (VGP_OPT_DIST_EMULATE_MIN) at n = 50 000.
mkdir -p gpurun_out/r02u
O=gpurun_out/r02u
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    tools/dist_inverse_bench.py 50000 dist_emulate_min=-1 dist_emulate_min=2048 dist_emulate_min=8192 \
    dist_emulate_min=16384 gemm_emulate_slices=0 > $O/dist_inverse_bench.jsonl 2> $O/err.txt
echo "rc=$?" | tee $O/rc.txt
tail -c 500 $O/err.txt
cat $O/dist_inverse_bench.jsonl | cut -c1-400
