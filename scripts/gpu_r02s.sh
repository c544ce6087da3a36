#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -8
bash scripts/gpu_ncu_kernel.sh r02s_describe_b "k_describe_b" 0 1 16 noise
