#!/bin/bash
# usage: tools/gpu_retry.sh <gpurun args...>  — re-submits while the pod answers "busy" (exit 3: nothing charged)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
