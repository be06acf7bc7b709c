#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...] : retries while the pod answers busy (exit 3), up to ~40 min
LOG=$1; shift
for i in $(seq 1 16); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
