#!/bin/bash
# gpurun_retry.sh LOG [gpurun args...]: retries while the pod answers "no box / slot free" (exit 3)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] attempt $i rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "[retry] gave up" >> "$log"; exit 3
