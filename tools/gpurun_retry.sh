#!/bin/bash
# Retries a gpurun call while the pod answers "transient" (nothing charged); usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "[gpurun_retry] gave up"
