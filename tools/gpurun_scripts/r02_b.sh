#!/bin/bash
# Round 2, call B: full GPU suite with the int8 products on by default + config-size parity tests, then both bench arms
# exactly as the driver runs them.
mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider --durations=15 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -40 $O/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/rc.txt
tail -c 3000 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b/bench_n1.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "e2e", "parity", "gpu_launches", "setup_s", "clocks"):
    print(k, json.dumps(d.get(k))[:1500])
print("cpu", json.dumps({k: v for k, v in d.get("cpu_baseline", {}).items() if k != "sample"})[:1500])
print("elbo", json.dumps(d.get("elbo"))[:1200])
print("lazy", json.dumps(d.get("lazy_column"))[:1500])
PY
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
echo "ref rc=$?" | tee -a $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b/bench_ref.json").read().strip().splitlines()[-1])
print("ref value", d["value"], "e2e", d["e2e"], "measured_at", d["measured_at"])
print(json.dumps(d["cpu_baseline"]["extrapolation"]))
PY
