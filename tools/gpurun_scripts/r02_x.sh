#!/bin/bash
# Round 2, call X (1 GPU): the whole GPU suite twice (fresh process each) after moving the score-buffer allocation out
# of the rank threads of the single-device emulation.
mkdir -p gpurun_out/r02x
O=gpurun_out/r02x
for i in 1 2; do
timeout 400 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider --durations=3 > $O/pytest_gpu_$i.log 2>&1
echo "pytest $i rc=$?" | tee -a $O/rc.txt
tail -7 $O/pytest_gpu_$i.log | cut -c1-200
done
