#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers busy (exit code 3 / transient)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
cat /tmp/gpurun_last.txt
exit $rc
