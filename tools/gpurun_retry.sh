#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): tools/gpurun_retry.sh <timeout> <command...>
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
