"""Measure the FP64 roofline denominator on the box: cuBLAS DGEMM 8192^3 via torch.matmul (burst: best of 10;
sustained: back to back for 3 s), plus our own DMMA GEMM on the same shape through vgp_dgemm's padded fast path.
Writes gpurun_out/fp64_peak.json.  SURVEY.md section 7.2: MEASURED_PEAKS.json has no FP64 figure."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi  # noqa: E402

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
flop = 2.0 * n ** 3


def timed(fn, reps):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
burst = flop / (timed(lambda: torch.matmul(a, b, out=c), 10) * 1e-3) / 1e12
t0 = time.time()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
reps = 0
while time.time() - t0 < 3.0:
    torch.matmul(a, b, out=c)
    reps += 1
    if reps % 8 == 0:
        torch.cuda.synchronize()
e1.record()
e1.synchronize()
sustained = flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12

stream = torch.cuda.current_stream().cuda_stream
out = {"cublas_dgemm_burst_tflops": burst, "cublas_dgemm_sustained_tflops": sustained, "n": n}
for name, (ta, tb) in {"nn": (0, 0), "nt": (0, 1), "tn": (1, 0)}.items():
    def ours():
        _ffi.call("vgp_dgemm", 0, ta, tb, n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n,
                  stream)
    for _ in range(2):
        ours()
    torch.cuda.synchronize()
    out["ours_dmma_%s_tflops" % name] = flop / (timed(ours, 5) * 1e-3) / 1e12
ref = torch.matmul(a, b)
_ffi.call("vgp_dgemm", 0, 0, 0, n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n, stream)
torch.cuda.synchronize()
out["ours_vs_cublas_max_rel_err"] = float(((c - ref).abs().max() / ref.abs().max()).item())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
