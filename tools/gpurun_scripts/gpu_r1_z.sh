#!/bin/bash
mkdir -p gpurun_out
export VGP_GEMM_CFG=tma
timeout 600 python -m pytest tests/test_gpu_expquad_dense.py tests/test_gpu_dist_inverse.py tests/test_gpu_gp.py tests/test_gpu_elbo.py tests/test_gpu_lazy.py -m gpu -q --maxfail=5 --timeout 120 -p no:cacheprovider > gpurun_out/pytest_tma.log 2>&1
echo "pytest tma exit $?"; tail -6 gpurun_out/pytest_tma.log | cut -c1-300
cat > /tmp/sweep_nt.py <<'PY'
import os, sys, torch, json
sys.path.insert(0, os.getcwd())
from vgposp_b200 import _ffi
stream = torch.cuda.current_stream().cuda_stream
for (m, n, k) in [(1024,1024,1024),(4096,4096,4096),(8192,8192,8192),(16384,16384,16384),(24960,24960,2048),(24960,512,512),(24960,128,128),(512,512,200064)]:
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(n, k, dtype=torch.float64, device="cuda"); c = torch.empty(m, n, dtype=torch.float64, device="cuda")
    def ours(): _ffi.call("vgp_dgemm", 0, 0, 1, m, n, k, 1.0, a.data_ptr(), k, b.data_ptr(), k, 0.0, c.data_ptr(), n, stream)
    for _ in range(2): ours()
    torch.cuda.synchronize()
    ref = a @ b.t(); err = float((c - ref).abs().max() / ref.abs().max())
    reps = max(3, min(50, int(2e12 / (2.0*m*n*k))))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): ours()
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1) / reps
    print(os.environ.get("VGP_GEMM_CFG"), m, n, k, "%.2f TFLOP/s" % (2.0*m*n*k/(ms*1e-3)/1e12), "err %.1e" % err, flush=True)
    del a, b, c, ref
PY
timeout 600 python /tmp/sweep_nt.py 2>&1 | tail -9
VGP_GEMM_CFG=pair timeout 600 python /tmp/sweep_nt.py 2>&1 | tail -9
timeout 600 python tools/e2e_only.py 3 auto 2>&1 | grep overlap | sed 's/^/tma /'
VGP_GEMM_CFG=pair timeout 600 python tools/e2e_only.py 3 auto 2>&1 | grep overlap | sed 's/^/pair /'
