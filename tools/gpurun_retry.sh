#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...> : retries while the pod answers "busy / draining" (exit 3, nothing charged)
log=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
