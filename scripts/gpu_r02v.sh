#!/bin/bash
# evidence run: default bench line, reference arm, configs[4] stream line, integer pipeline, launch list, ncu of the changed kernels
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r02v_bench_ours.json 2> gpurun_out/r02v_bench_ours.err; tail -c 300 gpurun_out/r02v_bench_ours.json; tail -3 gpurun_out/r02v_bench_ours.err
timeout 500 python bench.py --impl reference > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; tail -c 300 gpurun_out/r02v_bench_ref.json; tail -3 gpurun_out/r02v_bench_ref.err
timeout 300 python bench.py --config stream > gpurun_out/r02v_bench_stream.json 2> gpurun_out/r02v_bench_stream.err; tail -c 400 gpurun_out/r02v_bench_stream.json; tail -3 gpurun_out/r02v_bench_stream.err
timeout 300 python bench_fast.py 2>/dev/null | tail -2 > gpurun_out/r02v_bench_fast.json; cut -c1-300 gpurun_out/r02v_bench_fast.json
bash scripts/gpu_launches.sh r02v
