#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <gpus> <command...>   -- retries while the pod answers "busy" (exit 3)
T=$1; G=$2; shift 2
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
