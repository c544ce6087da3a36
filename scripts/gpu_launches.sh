#!/bin/bash
# launch list (device time per launch, serialised, cold cache) of one chunk of 32 frames on one lane
TAG=$1
mkdir -p gpurun_out
CMD="python bench.py --frames 32 --chunk 32 --lanes 1 --steps 1 --warmup 1 --no-cpu --no-match"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "rc=$?"
