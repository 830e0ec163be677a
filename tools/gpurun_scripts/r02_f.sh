#!/bin/bash
# Round 2, call F (2 GPUs): whole GPU suite (2-process IPC tests included), then the N=2 bench as the driver runs it.
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=12 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -45 $O/pytest_gpu.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench2 rc=$?" | tee -a $O/rc.txt
tail -c 1500 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches", "setup_s", "clocks"):
    print(k, json.dumps(d.get(k))[:1800])
print("elbo", json.dumps(d.get("elbo"))[:800])
PY
