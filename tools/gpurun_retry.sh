#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
#   tools/gpurun_retry.sh <log> <gpurun args...>
LOG="$1"; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
