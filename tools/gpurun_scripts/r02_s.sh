#!/bin/bash
mkdir -p gpurun_out/r02s
timeout 600 python -m pytest tests/test_gpu_dist_inverse.py -m gpu -q --timeout 300 -p no:cacheprovider --durations=8 > gpurun_out/r02s/pytest.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02s/rc.txt
grep -E "^E  |passed|failed|s call" gpurun_out/r02s/pytest.log | cut -c1-300 | head -20
