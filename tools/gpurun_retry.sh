#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3): tools/gpurun_retry.sh <timeout> '<command>'
to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
