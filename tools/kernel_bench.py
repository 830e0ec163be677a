"""Kernel-matrix builder throughput: symmetric n x n and rectangular n x (n / G) panels, every kernel kind.
Prints GB/s written (8 n1 n2 bytes per launch) next to MEASURED_PEAKS.json's copy bandwidth."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi  # noqa: E402
from vgposp_b200._ffi import call  # noqa: E402

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.default_rng(1)
x = rng.uniform(-2, 2, (n, 3))
ls = 0.5 * (1000.0 / n) ** (1.0 / 3.0)
xd = _ffi.DeviceArray.from_host(x, 0)
ld = n + (n % 2)
out = _ffi.DeviceArray((n, ld), np.float64, 0)
res = []


def timed(fn, reps=5):
    fn()
    call("vgp_stream_sync", 0, None)
    best = 1e9
    for _ in range(reps):
        e0, e1, ms = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_float()
        call("vgp_event_record", 0, None, ctypes.byref(e0))
        fn()
        call("vgp_event_record", 0, None, ctypes.byref(e1))
        call("vgp_stream_sync", 0, None)
        call("vgp_event_elapsed_ms", 0, e0, e1, ctypes.byref(ms))
        best = min(best, ms.value)
    return best


for kind, name in enumerate(["expquad", "matern12", "matern32", "matern52"]):
    for g in (1, 2, 8):
        nloc = n // g
        c0 = nloc                       # a panel that is not the first: rectangular path
        if g == 1:
            fn = lambda: call("vgp_kernel_matrix", 0, kind, xd.ptr, n, xd.ptr, n, 3, 1.0, ls, 1e-2, 0, out.ptr, ld, None)
        else:
            fn = lambda: call("vgp_kernel_matrix", 0, kind, xd.ptr, n, xd.ptr + c0 * 24, nloc, 3, 1.0, ls, 1e-2, c0,
                              out.ptr, ld, None)
        ms = timed(fn)
        gbs = 8.0 * n * nloc / (ms * 1e-3) / 1e9
        rec = {"kind": name, "n": n, "cols": nloc, "path": "symmetric" if g == 1 else "rectangular", "ms": ms,
               "GBps": gbs, "frac_of_measured_hbm": gbs / peak}
        print(rec, flush=True)
        res.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/kernel_bench_n%d.json" % n, "w"), indent=1)
