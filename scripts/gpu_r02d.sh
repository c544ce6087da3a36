#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/fed_one.py 1920 1080 32 3"
$CMD > gpurun_out/plain_fed4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fed4 -s 1 -c 2 -f -o gpurun_out/prof_r02d_fed4 $CMD > gpurun_out/ncu_r02d_fed4.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_r02d_fed4.log
