#!/bin/bash
# re-entry baseline: parity suite + default bench line
set -o pipefail
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
timeout 400 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 6000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
