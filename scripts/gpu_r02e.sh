#!/bin/bash
echo "== default dispatch"; python scripts/fed_probe.py --check 2>&1 | tail -17
for b in 64 96; do echo "== band $b"; AKZ_FED_MIN_UNITS=0 AKZ_FED_BAND=$b python scripts/fed_probe.py 2>&1 | tail -17 | head -8; done
