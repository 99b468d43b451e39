#!/bin/bash
# gpurun with retries on "busy / transient" answers (exit 3 or status=transient): tools/gpurun_retry.sh [gpurun args...]
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box\|retry in a few minutes" || [ $rc -eq 3 ]; then
    echo "[retry $i] transient answer, sleeping 90 s" >&2; sleep 90; continue
  fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
