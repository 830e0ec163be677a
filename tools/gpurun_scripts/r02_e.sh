#!/bin/bash
mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
for extra in 1 3 5; do
  echo "== 1000 2 512 1 extra=$extra" | tee -a $O/debug.log
  timeout 150 python tools/debug_dist_emul.py 1000 2 512 1 $extra 2>&1 | tail -9 | cut -c1-400 | tee -a $O/debug.log
done
