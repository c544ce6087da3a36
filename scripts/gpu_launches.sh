#!/bin/bash
# launch list (device time per launch, serialised, cold cache) of one chunk of 32 frames on one lane: bash scripts/gpu_launches.sh TAG [content]
TAG=$1; CONTENT=${2:-shapes}
mkdir -p gpurun_out
CMD="python scripts/chunk_once.py 32 3 $CONTENT"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "rc=$?"
