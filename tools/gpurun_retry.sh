#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> <gpus> <command string>   -- retries while the pod answers "busy" (rc 3)
t=$1; n=$2; shift 2
for i in $(seq 1 40); do
  if [ "$n" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $t -- "$@"; else /usr/local/graft/bin/gpurun --gpus $n --timeout $t -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i; sleeping 90 s"; sleep 90
done
exit 3
