#!/bin/bash
# Round 2, call K (2 GPUs): whole suite (IPC tests run), N = 2 bench as the driver runs it, N = 1 ELBO timing.
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -8 $O/pytest_gpu.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench2 rc=$?" | tee -a $O/rc.txt
tail -c 600 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "gpu_launches", "setup_s"):
    print(k, json.dumps(d.get(k))[:900])
e = d.get("e2e") or {}
print("e2e", e.get("value"), json.dumps(e.get("seconds"))[:400])
el = d.get("elbo") or {}
print("elbo", el.get("value"), el.get("ms_per_step"), json.dumps(el.get("roofline"))[:300])
PY
timeout 300 python tools/elbo_steps.py --steps 6 2>&1 | tail -3 | tee $O/elbo_int8.log
timeout 200 python tools/emulated_gemm_bench.py 8192 8 2>&1 | tail -1 | tee $O/emu.log
