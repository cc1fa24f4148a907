#!/usr/bin/env bash
# Local helper: repeat a gpurun call while the pod answers "busy" (exit code 3, nothing charged).  usage: gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
