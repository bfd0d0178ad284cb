#!/bin/bash
# usage: tools/gpurun_retry.sh <tries> <gpurun args...>   — retries while gpurun answers "no slot" (exit 3)
tries=$1; shift
for i in $(seq 1 "$tries"); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
