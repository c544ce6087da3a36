#!/bin/bash
# full GPU suite, default bench line (with the per-octave roofline rows), ncu --set full of the three streaming scale-space kernels (final versions)
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 700 -p no:cacheprovider -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -6
timeout 500 python bench.py > gpurun_out/r02z_bench_ours.json 2> gpurun_out/r02z_bench_ours.err; tail -c 300 gpurun_out/r02z_bench_ours.json; tail -3 gpurun_out/r02z_bench_ours.err
bash scripts/gpu_ncu_kernel.sh r02z_fed4 "k_fed4<3" 0 2 32
bash scripts/gpu_ncu_kernel.sh r02z_blur4 "k_blur4<1" 0 2 32
bash scripts/gpu_ncu_kernel.sh r02z_deriv4 "k_deriv4<3" 0 2 32
