#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): nothing is charged for those.
# usage: tools/gpurun_retry.sh <timeout_s> <command...>
t=$1; shift
for attempt in $(seq 1 40); do
  out=$(gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 40 attempts"; exit 3
