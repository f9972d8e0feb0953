#!/bin/bash
# usage: tools/gpurun_retry.sh TIMEOUT 'command'  [extra gpurun args, e.g. --gpus 2]
# gpurun answers "transient" (nothing charged) when the pod has no free slot: wait and ask again.
t=$1; cmd=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t "$@" -- "$cmd" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|no box or slot"; then sleep 120; continue; fi
  break
done
