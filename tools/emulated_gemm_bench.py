"""Time the int8-tensor-core FP64 GEMM (vgp_gemm_emulated) against cuBLAS DGEMM on one shape.
    python tools/emulated_gemm_bench.py [n] [slices]
The emulated entry allocates and slices per call; the slicing share is reported by timing a k = 128 call."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = int(sys.argv[2]) if len(sys.argv) > 2 else 8
stream = torch.cuda.current_stream().cuda_stream
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty(n, n, dtype=torch.float64, device="cuda")
ref = torch.empty(n, n, dtype=torch.float64, device="cuda")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emu(k):
    _ffi.call("vgp_gemm_emulated", 0, 0, 1, n, n, k, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n, s, 0,
              stream)


ms_emu = timed(lambda: emu(n))
ms_small = timed(lambda: emu(128))
ms_blas = timed(lambda: torch.matmul(a, b.t(), out=ref))
emu(n)
torch.cuda.synchronize()
err = float((c - ref).abs().max() / ref.abs().max())
flop = 2.0 * n ** 3
print(json.dumps({"n": n, "slices": s, "int8_products": s * (s + 1) // 2, "emulated_ms": ms_emu, "k128_call_ms": ms_small,
                  "cublas_dgemm_ms": ms_blas, "emulated_fp64_equiv_tflops": flop / ms_emu / 1e9,
                  "cublas_tflops": flop / ms_blas / 1e9, "int8_tops": s * (s + 1) / 2 * flop / ms_emu / 1e9,
                  "max_rel_diff_vs_cublas": err}))
