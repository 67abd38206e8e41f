#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> <command...>  -- retries while the pod answers "busy" (rc 3 / transient)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy" $LOG; then sleep 90; continue; fi
  break
done
