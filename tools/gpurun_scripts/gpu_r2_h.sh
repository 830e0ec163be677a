#!/bin/bash
VGP_LIB=$PWD/vgposp_b200/lib/libvgposp_head.so timeout 300 python tools/elbo_determinism.py 2>&1 | grep "first call" | sed 's/^/HEAD /'
timeout 300 python tools/elbo_determinism.py 2>&1 | grep "first call" | sed 's/^/NEW  /'
VGP_ELBO_OVERLAP=28 timeout 300 python tools/elbo_determinism.py 2>&1 | grep "first call" | sed 's/^/NEW28 /'
