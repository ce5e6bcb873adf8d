#!/bin/bash
# usage: gpurun_retry.sh <script> <log> <timeout_s> [gpus]   -- retries while the pod answers "transient"/busy (nothing charged)
S=$1; L=$2; T=$3; G=${4:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > $L 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > $L 2>&1; fi
  if grep -q "status=transient\|no box\|exit code 3\|rc=3" $L; then echo "retry $i $(date -u +%T)" >> $L.retries; sleep 150; continue; fi
  break
done
echo done >> $L.retries
