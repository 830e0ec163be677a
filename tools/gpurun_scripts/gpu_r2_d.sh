#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 120 -p no:cacheprovider > gpurun_out/pytest_small.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_small.log | cut -c1-300
cat > /tmp/sweep_small.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vgposp_b200 import _ffi
stream = torch.cuda.current_stream().cuda_stream
for (m, n, k) in [(512,512,512),(1024,1024,1024),(128,4096,128),(6144,128,128),(6144,512,512),(1024,256,1024),(2048,2048,2048)]:
  for tb in (0,1):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn((n,k) if tb else (k,n), dtype=torch.float64, device="cuda"); c = torch.empty(m, n, dtype=torch.float64, device="cuda")
    def ours(): _ffi.call("vgp_dgemm", 0, 0, tb, m, n, k, 1.0, a.data_ptr(), k, b.data_ptr(), b.shape[1], 0.0, c.data_ptr(), n, stream)
    for _ in range(2): ours()
    torch.cuda.synchronize()
    ref = a @ (b.t() if tb else b); err = float((c - ref).abs().max() / ref.abs().max())
    reps = 50
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): ours()
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1) / reps
    print(os.environ.get("VGP_GEMM_SMALL_BELOW","74"), m, n, k, "tb", tb, "%.1f us  %.2f TFLOP/s" % (ms*1e3, 2.0*m*n*k/(ms*1e-3)/1e12), "err %.1e" % err, flush=True)
PY
timeout 300 python /tmp/sweep_small.py 2>&1 | tail -14
VGP_GEMM_SMALL_BELOW=0 timeout 300 python /tmp/sweep_small.py 2>&1 | tail -14
timeout 600 python tools/e2e_only.py 3 auto 2>&1 | grep overlap
VGP_GEMM_SMALL_BELOW=0 timeout 600 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | sed 's/^/big-only /'
timeout 600 python tools/elbo_profile.py 5 2>&1 | tail -2
VGP_GEMM_SMALL_BELOW=0 timeout 600 python tools/elbo_profile.py 5 2>&1 | tail -2 | sed 's/^/big-only /'
