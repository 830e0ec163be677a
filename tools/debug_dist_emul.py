"""Debug helper (GPU): the sharded lazy-factor test as a script with knobs.
    python tools/debug_dist_emul.py <n> <world> <emulate_min> <poison 0|1>"""
import ctypes
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi, greedy  # noqa: E402
from vgposp_b200.dist_inverse import DistInverse, UPLOAD_LOWER  # noqa: E402

n, world, emin, poison = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
extra = int(sys.argv[5]) if len(sys.argv) > 5 else 0     # 1: lazy handles exist, 2: load_cov_device, 4: record_scores
D, k = 0, 12
_ffi.set_option("dist_min_tiles", 2)
_ffi.set_option("dist_min_k", 256)
_ffi.set_option("gemm_emulate_min", emin)
rng = np.random.default_rng(n + 1)
x = rng.uniform(-2, 2, (n, 3))
d = x[:, None, :] - x[None, :, :]
ls = 0.5 * (1000.0 / n) ** (1 / 3)
a = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)
want = greedy.place_single(a, k, D, want_step_scores=True, formulation="lazy_factor")
print("single ok", want[0][:5], flush=True)
streams = []
for _ in range(world):
    s = ctypes.c_void_p()
    _ffi.call("vgp_stream_create", D, ctypes.byref(s))
    streams.append(s)
ranks = [DistInverse(n, r, world, D, stream=streams[r]) for r in range(world)]
for r in ranks:
    r.connect_pointers([q.pointers for q in ranks])
    r.fill_padding()
    if poison:
        r.load_host(np.full((n, n), np.nan))
lazies = [greedy.LazyGreedy.from_dist(r, k) for r in ranks] if extra & 1 else None
bounds = [(n * g) // world for g in range(world + 1)]
for r in ranks:
    r0, r1 = bounds[r.rank], bounds[r.rank + 1]
    r.upload_rows(a[r0:r1], r0, r1, ncols=UPLOAD_LOWER)
errors = [None] * world
times = [None] * world


def work(i):
    t0 = time.perf_counter()
    try:
        r = ranks[i]
        r.barrier()
        if extra & 2:
            lazies[i].load_cov_device(r.ptr, r.ld)
        if extra & 4:
            lazies[i].record_scores(True)
        r.factor_inverse()
    except Exception as e:       # noqa: BLE001
        errors[i] = repr(e)
    times[i] = time.perf_counter() - t0


threads = [threading.Thread(target=work, args=(i,)) for i in range(world)]
for t in threads:
    t.start()
for t in threads:
    t.join(timeout=300)
print("errors", errors, "times", times, flush=True)
for r in ranks:
    print("rank", r.rank, r.stats(), flush=True)
got = [np.tril(r.to_host()) for r in ranks]
ref = np.tril(np.linalg.inv(np.linalg.cholesky(a)))
for i, g in enumerate(got):
    bad = ~np.isfinite(g)
    print("rank", i, "nonfinite", int(bad.sum()), "max abs diff vs numpy L^-1", float(np.nanmax(np.abs(g - ref))),
          "equal rank0", bool(np.array_equal(g, got[0])), flush=True)
