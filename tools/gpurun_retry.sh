#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "transient / busy" (nothing charged)
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3"; then sleep 90; continue; fi
  break
done
