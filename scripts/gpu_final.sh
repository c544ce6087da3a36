#!/bin/bash
# final evidence of a round: default bench lines (both arms), launch list + full captures of the top kernels, matcher capture
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default_ours.json 2> gpurun_out/bench_default_ours.err; tail -c 600 gpurun_out/bench_default_ours.json; tail -3 gpurun_out/bench_default_ours.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; tail -c 300 gpurun_out/bench_default_ref.json; tail -3 gpurun_out/bench_default_ref.err
timeout 700 bash scripts/gpu_ncu.sh $1 k_prep2 k_fed3
python scripts/match_probe.py && ncu --set full --clock-control none --import-source on -k regex:k_match -c 2 -f -o gpurun_out/prof_$1_k_match python scripts/match_probe.py > gpurun_out/ncu_match_$1.log 2>&1
echo "match ncu rc=$?"
