#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (nothing is charged for those).
#   tools/gpurun_retry.sh [gpurun options] -- '<command>'
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box or slot"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
