"""Drive `downdate_kernel` on ONE shard of a G-way split (columns [0, n / G) of the panel) in a single process, for an
`ncu --set full` capture of its DRAM traffic at the shard sizes the multi-GPU bench lines run (ncu is single-GPU only).
The panel values are synthetic (Sigma panel from the workload's cloud, identity precision): traffic does not depend
on them.  Usage: python tools/downdate_traffic.py n G [n G ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vgposp_b200 import _ffi, greedy  # noqa: E402
from vgposp_b200._ffi import call  # noqa: E402

args = [int(v) for v in sys.argv[1:]] or [50000, 2]
for n, g in zip(args[::2], args[1::2]):
    x, amp, ls, nugget = bench.workload(n)
    xd = _ffi.DeviceArray.from_host(x, 0)
    bounds = greedy.shard_bounds(n, g)
    nloc = bounds[1]
    shard = greedy.GreedyShard(n, 0, nloc, 4, 0)
    shard.build_cov_expquad(xd.ptr, 3, amp, ls, nugget)
    call("vgp_memset", 0, shard.prec_ptr, 0, shard.n_pad * shard.ld * 8, None)
    ones = np.ones(nloc)
    call("vgp_memcpy2d_h2d", 0, shard.prec_ptr, (shard.ld + 1) * 8, ones.ctypes.data, 8, 8, nloc, None)
    shard.reset()
    stride = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
    stride += stride % 2
    rec = _ffi.DeviceArray((4 * g,), np.float64, 0).zero_()
    seg = _ffi.DeviceArray((2 * stride,), np.float64, 0).zero_()
    segs = _ffi.DeviceArray((2 * stride * g,), np.float64, 0).zero_()
    for step in range(2):
        shard.local_best(rec)
        shard.select(rec, 1)
        shard.segments(seg, stride)
        for q in range(g):      # every "rank" contributes a copy of this shard's segment: right shape, any values
            call("vgp_memcpy_d2d", 0, segs.ptr + q * 2 * stride * 8, seg.ptr, 2 * stride * 8, None)
        shard.apply(segs, stride, bounds)
    shard.sync()
    sel, _ = shard.results()
    print("n", n, "G", g, "nloc", nloc, "selections", [int(v) for v in sel], flush=True)
    shard.close()
    xd.free()
