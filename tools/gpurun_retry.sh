#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / status=transient): tools/gpurun_retry.sh [gpurun args...]
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then echo "[retry] attempt $attempt: busy, sleeping"; sleep 150; continue; fi
  exit $rc
done
exit 3
