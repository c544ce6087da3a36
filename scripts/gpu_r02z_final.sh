#!/bin/bash
# evidence run: default bench line, reference arm, configs[4] stream line, integer pipeline, launch list, ncu of the changed kernels
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r02z_bench_ours.json 2> gpurun_out/r02z_bench_ours.err; tail -c 300 gpurun_out/r02z_bench_ours.json; tail -3 gpurun_out/r02z_bench_ours.err
timeout 500 python bench.py --impl reference > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; tail -c 300 gpurun_out/r02z_bench_ref.json; tail -3 gpurun_out/r02z_bench_ref.err
timeout 300 python bench.py --config stream > gpurun_out/r02z_bench_stream.json 2> gpurun_out/r02z_bench_stream.err; tail -c 400 gpurun_out/r02z_bench_stream.json; tail -3 gpurun_out/r02z_bench_stream.err
timeout 300 python bench_fast.py 2>/dev/null | tail -2 > gpurun_out/r02z_bench_fast.json; cut -c1-300 gpurun_out/r02z_bench_fast.json
bash scripts/gpu_launches.sh r02z
python scripts/chunk_once.py 1 2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02z_single.csv python scripts/chunk_once.py 1 3 > gpurun_out/ncu_r02z_single.log 2>&1
python scripts/single_frame_probe.py 2>&1 | tail -4 > gpurun_out/r02z_single_frame_probe.txt; cat gpurun_out/r02z_single_frame_probe.txt
