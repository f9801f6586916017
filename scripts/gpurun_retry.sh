#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers busy (nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|rc=3\|retry in a few minutes" $log; then break; fi
  sleep 90
done
