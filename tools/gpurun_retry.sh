#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] busy (attempt $i), retrying in 120 s" >&2
    sleep 120
done
exit 3
