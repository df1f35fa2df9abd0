#!/bin/bash
# Retries `gpurun` while the pod answers "busy" (exit code 3: nothing charged).  Usage: gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] pod busy, sleeping 90 s" >&2
  sleep 90
done
exit 3
