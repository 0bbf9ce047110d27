#!/bin/bash
# gpu_retry.sh TIMEOUT LOG -- CMD : runs gpurun, retrying while the pod answers "busy" (exit code 3)
TO=$1; LOG=$2; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
