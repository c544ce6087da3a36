#!/bin/bash
# descriptor parity tests + short bench line
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -x -k "desc or 1088 or dropin or fast or golden or pipeline" 2>&1 | tail -5
bash scripts/quick_bench.sh
