"""One DMMA GEMM (ours) and one cuBLAS DGEMM of the same shape, for an `ncu --set full` side-by-side capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgposp_b200 import _ffi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
stream = torch.cuda.current_stream().cuda_stream
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    _ffi.call("vgp_dgemm", 0, 0, tb, n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n, stream)
    torch.matmul(a, b.t() if tb else b, out=c)
torch.cuda.synchronize()
print("done")
