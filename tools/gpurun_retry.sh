#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] -- <command>
# retries while the pod answers "transient / busy" (nothing is charged for those)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if grep -q "status=transient\|no box\|busy" $log && ! grep -q "status=ok" $log; then
    sleep 90
  else
    break
  fi
done
