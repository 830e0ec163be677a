#!/bin/bash
# Round 2, call T (8 GPUs): the N = 8 bench line as the driver runs it, then the same setup with FP64-pipe products
# (gemm_emulate_slices=0) to compare the distributed factorisation.
mkdir -p gpurun_out/r02t
O=gpurun_out/r02t
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench8 rc=$?" | tee $O/rc.txt
tail -c 600 $O/bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 8 --steps 20 --warmup 5 --option gemm_emulate_slices=0 --no-cfg5 --no-elbo --no-cpu --no-lazy \
    > $O/bench_n8_fp64pipe.json 2> $O/bench_n8_fp64pipe.err
echo "bench8 fp64 pipe rc=$?" | tee -a $O/rc.txt
tail -c 400 $O/bench_n8_fp64pipe.err
python - <<'PY'
import json
for name in ("bench_n8", "bench_n8_fp64pipe"):
    try:
        d = json.loads(open("gpurun_out/r02t/%s.json" % name).read().strip().splitlines()[-1])
    except Exception as e:
        print(name, "unreadable", e)
        continue
    print("==", name)
    for k in ("value", "ms_per_step", "gpu_launches", "setup_s"):
        print(k, json.dumps(d.get(k))[:900])
    r = d.get("roofline") or {}
    print("roofline frac", r.get("frac"), "whole", r.get("whole_step_frac"), json.dumps(r.get("per_rank")))
    e = d.get("e2e") or {}
    print("e2e", e.get("value"), json.dumps(e.get("seconds"))[:400])
    el = d.get("elbo") or {}
    print("elbo", el.get("value"), el.get("ms_per_step"))
    print("cfg5", json.dumps(d.get("cfg5_n100k"))[:1200])
PY
