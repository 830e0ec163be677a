"""Run a few VGP ELBO training steps at BASELINE configs[2] (for ncu launch lists / timing)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vgposp_b200.gp_functions as gpf  # noqa: E402
from vgposp_b200 import _ffi  # noqa: E402

print("options", _ffi.apply_env_options(), flush=True)      # VGP_OPT_<NAME>=<int>: this tool only, not the library

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--m", type=int, default=512)
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--steps", type=int, default=6)
args = ap.parse_args()
rng = np.random.default_rng(1)
x = rng.uniform(-2, 2, (args.n, 3))
y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(args.n)
z = rng.uniform(-2, 2, (args.m, 3))
tr = gpf.VgpTrainer(x, y, z, args.b)
for it in range(args.steps):
    idx = rng.integers(args.n, size=args.b)
    t0 = time.perf_counter()
    loss = tr.step(x[idx], y[idx])
    print("step", it, "loss", loss, "ms", (time.perf_counter() - t0) * 1e3, flush=True)
tr.close()
