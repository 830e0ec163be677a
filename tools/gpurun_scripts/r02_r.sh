#!/bin/bash
mkdir -p gpurun_out/r02r
timeout 600 python -m pytest tests/test_gpu_dist_inverse.py -m gpu -q --timeout 300 -p no:cacheprovider -k "ranks_as_threads" > gpurun_out/r02r/pytest.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02r/rc.txt
grep -E "^E  |passed|failed" gpurun_out/r02r/pytest.log | cut -c1-400 | head -12
timeout 300 python -m pytest tests/test_gpu_dist_inverse.py -m gpu -q --timeout 200 -p no:cacheprovider -k "ranks_as_threads and 1664 and int8" > gpurun_out/r02r/pytest_single.log 2>&1
echo "single rc=$?" | tee -a gpurun_out/r02r/rc.txt
grep -E "^E  |passed|failed" gpurun_out/r02r/pytest_single.log | cut -c1-400 | head -8
