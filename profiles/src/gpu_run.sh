#!/bin/bash
# usage: gpu_run.sh <timeout> <logname> <command...>; retries while the pod is busy
TO=$1; LOG=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > gpurun_out/$LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$LOG; then break; fi
  sleep 90
done
echo finished rc=$rc
