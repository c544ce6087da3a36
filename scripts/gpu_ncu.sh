#!/bin/bash
# ncu evidence for one round: launch list (device time per launch) + full captures of the top kernels.
# usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [<kernel-regex> ...]
TAG=$1; shift
CMD="python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu --no-match"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in "$@"; do
  N=$(echo $K | tr -cd 'a-zA-Z0-9_')
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 3 -f -o gpurun_out/prof_${TAG}_${N} $CMD > gpurun_out/ncu_full_${TAG}_${N}.log 2>&1
  echo "full $K rc=$?"
done
ls -la gpurun_out | tail -12
