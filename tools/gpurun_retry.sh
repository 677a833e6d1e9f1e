#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  sleep 90
done
