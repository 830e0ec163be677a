#!/bin/bash
# flakiness hunt on the final build: the GPU suite three times, the one-call path eight times, ELBO determinism
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -1 | cut -c1-120; done
timeout 900 python tools/e2e_only.py 8 auto 2>&1 | grep overlap | awk '{print $NF, $(NF-1), $(NF-2), $(NF-3), $(NF-4), $11, $12}' | sort | uniq -c
for i in 1 2 3 4 5; do timeout 300 python tools/elbo_determinism.py 2>&1 | grep -c "rows differing [1-9]" | sed "s/^/elbo bad calls: /"; done
