#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> [extra gpurun args]: retries while the pod answers "busy" (exit 3)
t=$1; shift; sc=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" --timeout $t -- "bash $sc" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
tail -40 /tmp/gpurun_last.log
echo "gpurun rc=$rc after $i tries"
