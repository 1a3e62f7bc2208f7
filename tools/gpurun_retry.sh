#!/bin/bash
# Retries a gpurun call while the pod answers "transient" (busy, nothing charged). Usage: gpurun_retry.sh <timeout> <cmd>
T=$1; shift
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt was transient; sleeping 90 s"; sleep 90; continue
  fi
  break
done
