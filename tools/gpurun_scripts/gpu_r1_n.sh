#!/bin/bash
# GEMM tile configurations: parity under each, shape sweep under each, one-call e2e under each
mkdir -p gpurun_out
for cfg in base deep pair; do
  export VGP_GEMM_CFG=$cfg
  timeout 600 python -m pytest tests/test_gpu_expquad_dense.py tests/test_gpu_dist_inverse.py tests/test_gpu_gp.py -m gpu -q --maxfail=5 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gemm_$cfg.log 2>&1
  echo "pytest $cfg exit $?"; tail -4 gpurun_out/pytest_gemm_$cfg.log | cut -c1-300
  timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_$cfg.log 2>&1
  echo "sweep $cfg exit $?"; python - <<PY
import json
for r in json.load(open("gpurun_out/gemm_sweep.json")):
    print("$cfg %6d %6d %6d ours %6.2f cublas %6.2f" % (r["m"], r["n"], r["k"], r["ours_tflops"], r["cublas_tflops"]))
PY
  cp gpurun_out/gemm_sweep.json gpurun_out/gemm_sweep_$cfg.json
  timeout 600 python tools/e2e_only.py 2 auto 2>&1 | grep overlap | sed "s/^/$cfg /"
done
