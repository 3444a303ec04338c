#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit code 3); usage: gpu_retry.sh <timeout> '<command>' [extra gpurun args]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
