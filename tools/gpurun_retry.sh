#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): tools/gpurun_retry.sh <timeout> [--gpus N] <command...>
t=$1; shift
g=""
if [ "$1" = "--gpus" ]; then g="--gpus $2"; shift 2; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t $g -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
