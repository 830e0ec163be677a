#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_cov_producer.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15 | cut -c1-250
timeout 600 python examples/placement_pipeline.py --cover 10 --samples 12 --k 8 --steps 100 --out gpurun_out/demo 2>&1 | tail -4 | cut -c1-400
ls gpurun_out/demo
