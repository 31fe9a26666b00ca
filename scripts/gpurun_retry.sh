#!/bin/bash
# usage: scripts/gpurun_retry.sh TIMEOUT SCRIPT OUTFILE   -- retries while the pod answers busy (exit 3)
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" > $3 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
