#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 600 python -m pytest tests/test_gpu_dist_inverse.py tests/test_gpu_emulated_gemm.py tests/test_gpu_runtime.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02q/pytest.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02q/rc.txt
tail -5 gpurun_out/r02q/pytest.log | cut -c1-250
