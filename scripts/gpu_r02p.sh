#!/bin/bash
echo "== cp.async (production)"; python scripts/fed_probe.py --check 2>&1 | tail -17 | head -8
echo "== TMA bulk copies + mbarrier"; AKZ_FED_TMA=1 python scripts/fed_probe.py --check 2>&1 | tail -17 | head -8
AKZ_FED_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "fed_cycle or scale_space_fused or 1088" 2>&1 | tail -3
