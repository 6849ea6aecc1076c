#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers busy / transient
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3\b"; then echo "[retry $i] pod busy, sleeping 90 s"; sleep 90; continue; fi
  break
done
