#!/bin/bash
# dev aid: gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ "$rc" != "3" ] && exit $rc
  sleep 45
done
exit 3
