#!/bin/bash
mkdir -p gpurun_out
echo "== stream kernel (band 96)"; python scripts/fed_probe.py --check 2>&1 | tail -20
echo "== tile kernel"; AKZ_FED_STREAM=0 python scripts/fed_probe.py 2>&1 | tail -20
for b in 48 64 135; do echo "== band $b"; AKZ_FED_BAND=$b python scripts/fed_probe.py 2>&1 | tail -17; done
timeout 600 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -s -k "fed_cycle or dropin or cumatch or akazer or stage_functions or fed_tau" 2>&1 | tail -30
