#!/bin/bash
# scripts/gpurun_retry.sh TIMEOUT 'command'  -- retries while the pod answers busy (exit 3) or transient
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 45; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up"; exit 3
