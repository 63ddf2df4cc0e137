#!/bin/bash
# tools/gpurun_retry.sh <log> <gpurun args...> — retry while the pod has no free slot (exit code 3: nothing charged)
LOG=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
    sleep 90
done
exit 3
