#!/bin/bash
# 1-GPU call with retries while the pod answers "busy" (exit code 3); usage: tools/gpurun_retry.sh <timeout> <command...>
t=$1; shift
for k in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
