"""Where the `h2d_push` phase of ShardedPlacer.place goes (one GPU is enough for the host-side pieces):
chunked triangle upload of a row slab, the Sigma copy next to the factor (2-D and 1-D forms), the barrier.
Usage: python tools/upload_phase_bench.py [n] [world]"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi, greedy                     # noqa: E402
from vgposp_b200._ffi import call                        # noqa: E402
from vgposp_b200.dist_inverse import DistInverse         # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D = 0
inv = DistInverse(n, 0, 1, D)
inv.connect_pointers([inv.pointers])
inv.fill_padding()
lz = greedy.LazyGreedy.from_dist(inv, 4)
out = {"n": n, "world_emulated": world}
for g in (world - 1, 0):                                    # the last slab (widest rows) and the first
    b = greedy.triangle_bounds(n, world)
    r0, r1 = b[g], b[g + 1]
    rows = r1 - r0
    host = ctypes.c_void_p()
    call("vgp_host_alloc", rows * n * 8, ctypes.byref(host))
    slab = np.ctypeslib.as_array(ctypes.cast(host, ctypes.POINTER(ctypes.c_double)), shape=(rows, n))
    slab[:] = 1.0
    for rep in range(2):
        t = time.perf_counter()
        inv.upload_rows(slab, r0, r1, ncols=-1)
        call("vgp_stream_sync", D, None)
        dt = time.perf_counter() - t
    out["upload_slab_%d" % g] = {"rows": rows, "cols": r1, "GB": (r1 * r1 - r0 * r0) * 4 / 1e9, "s": dt,
                                 "GBps": (r1 * r1 - r0 * r0) * 4 / 1e9 / dt}
    t = time.perf_counter()
    call("vgp_memcpy2d_h2d", D, inv.ptr + r0 * inv.ld * 8, inv.ld * 8, slab.ctypes.data, n * 8, r1 * 8, rows, None)
    call("vgp_stream_sync", D, None)
    dt = time.perf_counter() - t
    out["one_2d_copy_slab_%d" % g] = {"s": dt, "GBps": (r1 * r1 - r0 * r0) * 4 / 1e9 / dt}
    call("vgp_host_free", host)
for rep in range(2):
    t = time.perf_counter()
    lz.load_cov_device(inv.ptr, inv.ld)
    call("vgp_stream_sync", D, None)
    out["cov_copy_2d_s"] = time.perf_counter() - t
    t = time.perf_counter()
    call("vgp_memcpy_d2d", D, lz.cov_ptr, inv.ptr, n * inv.ld * 8, None)
    call("vgp_stream_sync", D, None)
    out["cov_copy_1d_s"] = time.perf_counter() - t
    t = time.perf_counter()
    inv.barrier()
    call("vgp_stream_sync", D, None)
    out["barrier_s"] = time.perf_counter() - t
print(json.dumps(out))
