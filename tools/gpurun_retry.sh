#!/bin/bash
# gpurun with retries while the pod answers "transient"/busy (exit 3) -- nothing is charged for those.
# Usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$LOG" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
