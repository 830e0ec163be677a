#!/bin/bash
# Round 2, call M (2 GPUs): multi-GPU + int8 tests, N = 2 bench, ELBO and GEMM timings on one GPU.
mkdir -p gpurun_out/r02m
O=gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_dist_inverse.py tests/test_gpu_greedy.py tests/test_gpu_emulated_gemm.py tests/test_gpu_elbo.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -4 $O/pytest_gpu.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench2 rc=$?" | tee -a $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline"):
    print(k, json.dumps(d.get(k))[:1000])
e = d.get("e2e") or {}
print("e2e", e.get("value"), json.dumps(e.get("seconds"))[:400])
el = d.get("elbo") or {}
print("elbo", el.get("value"), el.get("ms_per_step"))
PY
timeout 300 python tools/elbo_steps.py --steps 6 2>&1 | tail -3 | tee $O/elbo_int8.log
for n in 4096 8192; do timeout 200 python tools/emulated_gemm_bench.py $n 8 2>&1 | tail -1 | tee -a $O/emu.log; done
