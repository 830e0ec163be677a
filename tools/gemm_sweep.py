"""DMMA GEMM efficiency sweep on the shapes the recursive factorisation issues (square, thin-k panels, tall trsm leaves),
next to cuBLAS (torch.matmul) on the same shapes.  Writes gpurun_out/gemm_sweep.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi  # noqa: E402

stream = torch.cuda.current_stream().cuda_stream
shapes = [(1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192), (16384, 16384, 16384),
          (24960, 128, 128), (24960, 256, 256), (24960, 512, 512), (24960, 1024, 1024), (24960, 2048, 2048),
          (12416, 128, 128), (12416, 512, 512), (6144, 128, 128), (6144, 512, 512),
          (24960, 24960, 2048), (512, 512, 200064), (512, 4096, 512)]
out = []
for (m, n, k) in shapes:
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    c = torch.empty(m, n, dtype=torch.float64, device="cuda")
    flop = 2.0 * m * n * k

    def ours():
        _ffi.call("vgp_dgemm", 0, 0, 0, m, n, k, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, stream)

    def cublas():
        torch.matmul(a, b, out=c)

    rec = {"m": m, "n": n, "k": k}
    for name, fn in (("ours", ours), ("cublas", cublas)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        reps = max(3, min(50, int(2e12 / flop)))
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec[name + "_ms"] = ms
        rec[name + "_tflops"] = flop / (ms * 1e-3) / 1e12
    print(rec, flush=True)
    out.append(rec)
    del a, b, c
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
