#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elbo.py tests/test_gpu_gp.py -m gpu -q --maxfail=5 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_elbo_overlap.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_elbo_overlap.log | cut -c1-300
timeout 600 python tools/elbo_profile.py 6 2>&1 | tail -3
compute-sanitizer --tool racecheck --print-limit 5 python - <<'PY' 2>&1 | tail -5
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import vgposp_b200.gp_functions as gpf
rng = np.random.default_rng(1); n, m, b = 3000, 128, 256
x = rng.uniform(-2, 2, (n, 3)); y = np.sum(np.sin(2*np.pi*x), axis=1); z = rng.uniform(-2, 2, (m, 3))
tr = gpf.VgpTrainer(x, y, z, b)
for i in range(2): print(tr.step(x[:b], y[:b]))
PY
