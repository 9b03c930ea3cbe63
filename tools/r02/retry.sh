#!/bin/bash
# usage: tools/r02/retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3)
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun exit $rc" >> "$out"; exit $rc; fi
  sleep 150
done
echo "gave up" >> "$out"
