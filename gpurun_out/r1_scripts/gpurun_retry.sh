#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "transient"/busy (exit 3)
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
