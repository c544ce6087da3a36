#!/bin/bash
# round 2, first GPU call: whole parity suite (new drop-in tests included), smoke, default bench line
set -o pipefail
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -s 2>&1 | tee gpurun_out/r02a_pytest_gpu.log | tail -60
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee gpurun_out/r02a_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_ours.json 2> gpurun_out/r02a_bench_ours.err; tail -c 3000 gpurun_out/r02a_bench_ours.json; tail -5 gpurun_out/r02a_bench_ours.err
