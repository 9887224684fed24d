#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>   - repeats the call while the pod answers "busy" (nothing charged)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
