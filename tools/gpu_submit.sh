#!/bin/bash
# usage: tools/gpu_submit.sh LOGFILE TIMEOUT 'command'   -- retries while the pod answers "busy" (nothing charged)
log=$1; to=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 45; continue; fi
  break
done
tail -5 "$log"
