#!/bin/bash
# Retry a gpurun call while the pod answers "busy / no slot" (exit code 3; nothing is charged).
#   tools/gpurun_retry.sh [gpurun options] -- '<command>'
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
