#!/bin/bash
# usage: scripts/gpucallN.sh <gpus> <name> <timeout> <script>   -- as gpucall.sh, on N GPUs of one box
n=$1; name=$2; tmo=$3; script=$4
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $tmo -- "bash $script" > gpurun_out/$name.out 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" gpurun_out/$name.out && ! grep -q "status=ok" gpurun_out/$name.out; then sleep 120; continue; fi
  break
done
tail -5 gpurun_out/$name.out
