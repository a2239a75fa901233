#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command' : retries while the pod answers "busy / transient" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
