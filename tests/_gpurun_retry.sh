#!/bin/bash
# usage: tests/_gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "busy", exit code 3)
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
