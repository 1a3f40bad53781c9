#!/bin/bash
# usage: scripts/gpucall.sh <name> <timeout> <script>   -- retries while the pod answers "busy" (nothing charged)
name=$1; tmo=$2; script=$3
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $tmo -- "bash $script" > gpurun_out/$name.out 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" gpurun_out/$name.out && ! grep -q "status=ok" gpurun_out/$name.out; then sleep 90; continue; fi
  break
done
tail -5 gpurun_out/$name.out
