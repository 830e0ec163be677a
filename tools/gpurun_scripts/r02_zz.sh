#!/bin/bash
# Round 2, last call (1 GPU, < 100 s): dense-greedy tests and the N = 1 resident line after the downdate grid change.
mkdir -p gpurun_out/r02zz
O=gpurun_out/r02zz
timeout 45 python -m pytest tests/test_gpu_greedy.py tests/test_gpu_baseline_configs.py -x -q -m gpu -p no:cacheprovider -k "not cfg1 and not cfg3" > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -3 $O/pytest.log | cut -c1-200
timeout 50 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-elbo --no-lazy --no-e2e > $O/bench_n1_resident.json 2> $O/bench.err
echo "bench rc=$?" | tee -a $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02zz/bench_n1_resident.json").read().strip().splitlines()[-1])
print(d.get("value"), d.get("ms_per_step"), json.dumps(d.get("roofline"))[:400])
PY
