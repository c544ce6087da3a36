#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -s 2>&1 | tee gpurun_out/r02h_pytest_gpu.log | grep -v "^\[1080p\] first" | tail -40
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3
