#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  Usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  echo "[retry] busy (attempt $i), sleeping 20 s"
  sleep 20
done
exit 3
