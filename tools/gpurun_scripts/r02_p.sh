#!/bin/bash
# Round 2, call N (8 GPUs): the N = 8 bench line as the driver runs it (includes the n = 100 000 extra).
mkdir -p gpurun_out/r02p
O=gpurun_out/r02p
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench8 rc=$?" | tee $O/rc.txt
tail -c 800 $O/bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p/bench_n8.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "gpu_launches", "setup_s"):
    print(k, json.dumps(d.get(k))[:1000])
e = d.get("e2e") or {}
print("e2e", e.get("value"), json.dumps(e.get("seconds"))[:500], "first", json.dumps(e.get("first_call"))[:300])
el = d.get("elbo") or {}
print("elbo", el.get("value"), el.get("ms_per_step"), json.dumps(el.get("roofline"))[:200])
print("cfg5", json.dumps(d.get("cfg5_n100k"))[:1500])
PY
