#!/bin/bash
# parity suite + the short bench line (classes of the default and the noise workload)
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
bash scripts/quick_bench.sh
