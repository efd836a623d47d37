#!/bin/bash
# usage: tools/gpr.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers busy (exit code 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpr_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpr_last.log; then sleep 90; continue; fi
  break
done
tail -100 /tmp/gpr_last.log
exit $rc
