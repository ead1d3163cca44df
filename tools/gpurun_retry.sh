#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3); usage: gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 120
done
tail -60 /tmp/gpurun_last.log
exit $rc
