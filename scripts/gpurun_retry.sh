#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.out 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.out | tail -100
exit $rc
