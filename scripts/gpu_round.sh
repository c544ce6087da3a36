#!/bin/bash
# one GPU call: parity suite, golden fixture, bench (both arms), ncu launch list and full captures of the two top kernels
set -o pipefail
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests -m gpu --tb=short --timeout 120 -p no:cacheprovider -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40
timeout 60 python tests/golden/make_ref_golden.py gpurun_out/ref_320x240.npz 2>&1 | tail -3
timeout 300 python bench.py --frames 64 --steps 3 --warmup 3 --cpu-frames 4 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
timeout 300 python bench.py --impl reference --frames 64 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 2000 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
