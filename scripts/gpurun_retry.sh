#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <command...>  — retries while the pod answers busy (exit 3 / "transient"), nothing is charged for those
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box or slot"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up"; exit 3
