#!/bin/bash
# usage: gpuretry.sh <logfile> <gpurun args...>   retries while the pod answers busy (exit 3 / transient)
LOG=$1; shift
for i in $(seq 1 12); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "gpuretry done rc=$rc" >> "$LOG"
