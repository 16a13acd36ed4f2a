#!/bin/bash
# usage: gpurun_retry.sh LOGFILE TIMEOUT_S [--gpus N] -- COMMAND     (retries while the pod answers "busy / draining": exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 75
done
exit 3
