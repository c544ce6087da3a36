#!/bin/bash
# full ncu capture of the first launches of one kernel in a chunk: bash scripts/gpu_ncu_kernel.sh TAG REGEX [skip] [count] [frames] [content]
TAG=$1; K=$2; SKIP=${3:-0}; CNT=${4:-2}; NF=${5:-32}; CONTENT=${6:-shapes}
mkdir -p gpurun_out
CMD="python scripts/chunk_once.py $NF 2 $CONTENT"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_${TAG}.log
