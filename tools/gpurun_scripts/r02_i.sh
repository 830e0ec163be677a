#!/bin/bash
# Round 2, call I (1 GPU): suite with the ELBO's N-sized products on the int8 path, bench as the driver runs it,
# emulate_min sweep of the one-call placement.
mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -8 $O/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i/bench_n1.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "e2e", "parity", "gpu_launches", "setup_s", "clocks"):
    print(k, json.dumps(d.get(k))[:1300])
print("elbo", json.dumps(d.get("elbo"))[:1500])
PY
for mn in 512 768; do
  VGP_OPT_GEMM_EMULATE_MIN=$mn timeout 300 python tools/e2e_only.py 2 auto 2>&1 | grep -E "overlap|options" | tee $O/e2e_min$mn.log
done
VGP_OPT_GEMM_EMULATE_SLICES=0 timeout 300 python tools/elbo_steps.py 2>&1 | tail -3 | tee $O/elbo_fp64.log
timeout 300 python tools/elbo_steps.py 2>&1 | tail -3 | tee $O/elbo_int8.log
