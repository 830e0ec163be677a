#!/bin/bash
mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
for args in "1000 2 2048 1" "1000 2 512 0" "1000 2 512 1"; do
  echo "== $args" | tee -a $O/debug.log
  timeout 400 python tools/debug_dist_emul.py $args 2>&1 | tail -12 | tee -a $O/debug.log
done
