#!/bin/bash
# k_level4: parity test, then the short bench line with and without it
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_parity.py -m gpu --tb=short --timeout 700 -p no:cacheprovider -q -x -k "level4" 2>&1 | tee gpurun_out/pytest_level4.log | tail -25
for k in "AKZ_LEVEL4=1" "AKZ_LEVEL4=1 AKZ_L4_CTAS=4" "AKZ_LEVEL4=0"; do echo "$k"; env $k bash scripts/quick_bench.sh; done 2>&1 | tee gpurun_out/level4_ab.txt
