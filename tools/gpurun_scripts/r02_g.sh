#!/bin/bash
# Round 2, call G (1 GPU): whole GPU suite after the pinv / Matern-ELBO work, then one ncu --set full capture of the int8
# tcgen05 GEMM kernel (plain run first, as the profiling recipe demands).
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -30 $O/pytest_gpu.log | cut -c1-250
timeout 300 python tools/emulated_gemm_bench.py 8192 8 > $O/emu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emu_gemm_resident -c 1 -o $O/emu_prof \
    python tools/emulated_gemm_bench.py 8192 8 > $O/emu_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/rc.txt
tail -3 $O/emu_plain.log | cut -c1-600
tail -5 $O/emu_ncu.log | cut -c1-300
