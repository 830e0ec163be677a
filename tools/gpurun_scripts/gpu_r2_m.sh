#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=5 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu9.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu9.log | cut -c1-200
cat > /tmp/sweep_t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vgposp_b200 import _ffi
stream = torch.cuda.current_stream().cuda_stream
for (m, n, k) in [(512,512,512),(4096,4096,4096),(8192,8192,8192),(16384,16384,16384)]:
  for ta, tb in ((1,0),(1,1)):
    a = torch.randn(k, m, dtype=torch.float64, device="cuda"); b = torch.randn((n,k) if tb else (k,n), dtype=torch.float64, device="cuda"); c = torch.empty(m, n, dtype=torch.float64, device="cuda")
    def ours(): _ffi.call("vgp_dgemm", 0, ta, tb, m, n, k, 1.0, a.data_ptr(), m, b.data_ptr(), b.shape[1], 0.0, c.data_ptr(), n, stream)
    for _ in range(2): ours()
    torch.cuda.synchronize()
    ref = a.t() @ (b.t() if tb else b); err = float((c - ref).abs().max() / ref.abs().max())
    reps = max(3, min(50, int(2e12 / (2.0*m*n*k))))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): ours()
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1) / reps
    print(os.environ.get("VGP_GEMM_CFG","tma"), m, n, k, "ta", ta, "tb", tb, "%.2f TFLOP/s" % (2.0*m*n*k/(ms*1e-3)/1e12), "err %.1e" % err, flush=True)
    del a, b, c, ref
PY
timeout 600 python /tmp/sweep_t.py 2>&1 | tail -8
VGP_GEMM_CFG=pair timeout 600 python /tmp/sweep_t.py 2>&1 | tail -8
timeout 600 python bench.py --no-e2e --no-cpu --no-elbo --no-lazy --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dense path:', d['value'], d['roofline']['frac'], json.dumps(d['setup_s']))"
