#!/bin/bash
# 2-GPU validation of the NCCL sharded greedy + GPU parity suite
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu4.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu4.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_n50k_g2.log 2>&1
echo "bench g2 exit $?"; tail -3 gpurun_out/bench_n50k_g2.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --impl reference > gpurun_out/bench_ref_g2.log 2>&1
echo "ref g2 exit $?"; tail -2 gpurun_out/bench_ref_g2.log | cut -c1-600
