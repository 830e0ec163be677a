#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=None" "$log" && ! grep -q "status=ok" "$log"; then
    sleep 150
    continue
  fi
  break
done
echo "attempts: $attempt" >> "$log"
