#!/bin/bash
for ov in 0 1 2 3 3; do VGP_ELBO_OVERLAP=$ov timeout 300 python tools/elbo_profile.py 6 2>&1 | tail -2 | sed "s/^/ov=$ov /"; done
