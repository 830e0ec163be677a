"""gpf.calc_H at the reference's size (160 x 160 grid of (length_scale, amplitude), 25 observations, main.py:400-419):
the one-launch batched sweep against the sequential per-point evaluation (timed on a 16 x 16 sub-grid)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vgposp_b200.gp_functions as gpf  # noqa: E402

rng = np.random.default_rng(3)
x = rng.uniform(-2, 2, (25, 2))
y = np.sin(x[:, 0]) * np.sin(x[:, 1]) + 0.1 + 0.1 * rng.standard_normal(25)
amp, amp_assign, amp_p, lensc, lensc_assign, lensc_p, _, _, _, noise = \
    gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.54]), np.array([0.3]))
gp = gpf.fit_gp(gpf.create_cov_kernel(amp, lensc), x, noise)
gpf.calc_H(4, 4, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, gp.log_prob, None, None, y)     # warm-up
t0 = time.perf_counter()
H = gpf.calc_H(160, 160, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, gp.log_prob, None, None, y)
batched = time.perf_counter() - t0
t0 = time.perf_counter()
seq = np.zeros((16, 16))
for i in range(16):
    for j in range(16):
        lensc_assign([40 * (1 + 10 * i) / 160])
        amp_assign([40 * (1 + 10 * j) / 160])
        seq[i, j] = gp.log_prob(y)
sequential = (time.perf_counter() - t0) / 256 * 25600
rec = {"grid": [160, 160], "n_obs": 25, "batched_s": batched, "evaluations_per_s": 25600 / batched,
       "sequential_s_extrapolated_from_256": sequential, "speedup": sequential / batched,
       "max_rel_diff_on_subgrid": float(np.max(np.abs(H[::10, ::10][:16, :16] - seq) / np.abs(seq)))}
print(json.dumps(rec))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open("gpurun_out/calc_h_bench.json", "w"), indent=1)
