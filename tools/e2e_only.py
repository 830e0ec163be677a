"""The one-call host path alone (vgp_placement_host_ex on a pinned host covariance), repeated; prints the event
breakdown of every call.  Options through the environment of this tool (VGP_OPT_<NAME>, _ffi.apply_env_options):
VGP_OPT_H2D_OVERLAP=0 finishes the copy before the factorisation starts; VGP_OPT_GEMM_EMULATE_SLICES=0 keeps every
product on the FP64 pipe."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vgposp_b200 import _ffi  # noqa: E402
from vgposp_b200._ffi import call  # noqa: E402
from vgposp_b200.greedy import FORMULATIONS  # noqa: E402

print("options", _ffi.apply_env_options(), flush=True)
n = int(os.environ.get("VGP_BENCH_N", 50000))
k = 100
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x, amp, ls, nugget = bench.workload(n)
xd = _ffi.DeviceArray.from_host(x, 0)
cov = _ffi.DeviceArray((n, n), np.float64, 0)
call("vgp_expquad_matrix", 0, xd.ptr, n, xd.ptr, n, 3, amp, ls, nugget, 0, cov.ptr, n, None)
host = ctypes.c_void_p()
call("vgp_host_alloc", 8 * n * n, ctypes.byref(host))
call("vgp_memcpy_d2h", 0, host, cov.ptr, 8 * n * n, None)
call("vgp_stream_sync", 0, None)
del cov
for form in sys.argv[2:] or ["auto"]:
    for r in range(reps):
        sel = np.full(k, -1, dtype=np.int64)
        sc = np.zeros(k)
        secs = np.zeros(4)
        t0 = time.perf_counter()
        call("vgp_placement_host_ex", 0, host, n, n, k, 1e-8, 0.0, FORMULATIONS[form], sel.ctypes.data, sc.ctypes.data,
             None, secs.ctypes.data)
        wall = time.perf_counter() - t0
        hw = np.zeros(4)
        call("vgp_placement_host_wall", hw.ctypes.data)
        print(form, "overlap=%d" % _ffi.get_option("h2d_overlap"), "h2d %.3f factor %.3f select %.3f total %.3f"
              % tuple(secs), "wall %.3f" % wall, "(alloc %.3f enqueue %.3f release %.3f)" % tuple(hw[:3]), "sel", sel[:4],
              flush=True)
call("vgp_host_free", host)
