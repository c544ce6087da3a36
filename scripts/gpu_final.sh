#!/bin/bash
# final evidence of a round: default bench lines (both arms), launch list + full captures of the top kernels, matcher captures
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default_ours.json 2> gpurun_out/bench_default_ours.err; tail -c 400 gpurun_out/bench_default_ours.json; tail -3 gpurun_out/bench_default_ours.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; tail -c 300 gpurun_out/bench_default_ref.json; tail -3 gpurun_out/bench_default_ref.err
timeout 700 bash scripts/gpu_ncu.sh $1 k_prep2 k_fed3
for K in 3; do
  python scripts/match_probe.py $K && ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/prof_$1_k_match_kernel$K python scripts/match_probe.py $K > gpurun_out/ncu_match_$1_$K.log 2>&1
  echo "match ncu $K rc=$?"
done
timeout 200 python bench_stream.py --frames 16 --chunk 8 --steps 2 2>/dev/null | tail -1 > gpurun_out/bench_stream.json; cat gpurun_out/bench_stream.json | cut -c1-300
timeout 200 python bench_fast.py 2>/dev/null | tail -2 > gpurun_out/bench_fast.json; cat gpurun_out/bench_fast.json | cut -c1-400
