#!/bin/bash
# Round 2, call Y (2 GPUs): final build -- whole GPU suite (two-process IPC tests included) and the N = 2 bench line.
mkdir -p gpurun_out/r02y
O=gpurun_out/r02y
timeout 400 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider --durations=3 > $O/pytest_gpu_2gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -7 $O/pytest_gpu_2gpu.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-lazy > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench2 rc=$?" | tee -a $O/rc.txt
tail -c 300 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02y/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches", "setup_s"):
    print(k, json.dumps(d.get(k))[:600])
r = d.get("roofline") or {}
print("roofline frac", r.get("frac"), "whole", r.get("whole_step_frac"), json.dumps(r.get("per_rank")))
e = d.get("e2e") or {}
print("e2e", e.get("value"), json.dumps(e.get("seconds"))[:400], e.get("selection_equals_resident_run"))
el = d.get("elbo") or {}
print("elbo", el.get("value"), el.get("ms_per_step"))
PY
