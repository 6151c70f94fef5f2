#!/bin/bash
# dev helper: gpurun with retries on "busy" (exit 3).  usage: tools/grun.sh LOG TIMEOUT 'command'
LOG="$1"; TMO="$2"; shift 2
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$TMO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 45
done
exit 3
