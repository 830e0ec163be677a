"""Breakdown of one VGP ELBO training step at BASELINE configs[2]: host wall vs device time (CUDA events)."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import vgposp_b200.gp_functions as gpf  # noqa: E402
from vgposp_b200._ffi import call  # noqa: E402

n, m, b = 200000, 512, 4096
rng = np.random.default_rng(1)
x = rng.uniform(-2, 2, (n, 3))
y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(n)
z = rng.uniform(-2, 2, (m, 3))
tr = gpf.VgpTrainer(x, y, z, b)
xt = torch.as_tensor(x, device="cuda:0")
yt = torch.as_tensor(y, device="cuda:0")
xb = torch.empty((b, 3), dtype=torch.float64, device="cuda:0")
yb = torch.empty((b,), dtype=torch.float64, device="cuda:0")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    t0 = time.perf_counter()
    idx = torch.as_tensor(rng.integers(n, size=b), device="cuda:0")
    xb.copy_(xt[idx])
    yb.copy_(yt[idx])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    call("vgp_event_record", 0, None, ctypes.byref(e0))
    loss = tr.step_device(xb.data_ptr(), yb.data_ptr())
    t2 = time.perf_counter()
    call("vgp_event_record", 0, None, ctypes.byref(e1))
    ms = ctypes.c_float()
    call("vgp_event_elapsed_ms", 0, e0, e1, ctypes.byref(ms))
    print("step %d loss %.6f  batch-gather host %.2f ms  step call host %.2f ms  device(events) %.2f ms"
          % (it, loss, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ms.value), flush=True)
tr.close()
