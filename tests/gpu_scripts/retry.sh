#!/bin/bash
# usage: tests/gpu_scripts/retry.sh <timeout_s> <gpus> <command...>   — retries while the pod answers busy (rc 3)
t=$1; g=$2; shift 2
for i in $(seq 1 40); do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $g --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?; fi
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 60
done
tail -25 /tmp/gpurun_last.log
exit $rc
