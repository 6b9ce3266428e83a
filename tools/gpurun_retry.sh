#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh [gpurun flags] -- '<command>'
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt answered busy; sleeping 150 s" >&2
  sleep 150
done
exit 3
