#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  Usage: gpurun_retry.sh OUTFILE [gpurun args...]
OUT=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  RC=$?
  if [ "$RC" != "3" ] && ! grep -q "status=transient" "$OUT"; then exit $RC; fi
  sleep 90
done
exit 3
