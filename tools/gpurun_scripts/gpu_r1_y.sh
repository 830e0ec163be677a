#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/downdate_traffic.py 50000 2 50000 4 50000 8 100000 8 > gpurun_out/downdate_traffic_plain.log 2>&1
echo "plain exit $?"; tail -5 gpurun_out/downdate_traffic_plain.log
timeout 900 ncu --set full --clock-control none -k regex:downdate -c 8 -f -o gpurun_out/prof_downdate_shards python tools/downdate_traffic.py 50000 2 50000 4 50000 8 100000 8 > gpurun_out/downdate_traffic_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_downdate_shards.ncu-rep --page raw --csv > gpurun_out/prof_downdate_shards_raw.csv 2>/dev/null
ls -la gpurun_out/prof_downdate_shards*
