#!/bin/bash
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
timeout 300 python -m pytest tests/test_gpu_dist_inverse.py -q -x --timeout 120 -p no:cacheprovider -k "sharded_lazy and 1000" > $O/a.log 2>&1
echo "a rc=$?"; tail -5 $O/a.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_dist_inverse.py -q --timeout 120 -p no:cacheprovider > $O/b.log 2>&1
echo "b rc=$?"; tail -15 $O/b.log | cut -c1-300
