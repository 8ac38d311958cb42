#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <out_file> <command...>   (retries while the pod has no free box)
T=$1; OUT=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$OUT" 2>&1
  rc=$?
  echo "exit $rc (attempt $i)" >> "$OUT"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
