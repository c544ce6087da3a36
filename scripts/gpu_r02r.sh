#!/bin/bash
# chunk-size sweep of the default workload + ncu of k_describe_s on the noise input
mkdir -p gpurun_out
for c in 32 64 128; do for l in 1 2; do
echo "== chunk $c lanes $l"; timeout 300 python bench.py --chunk $c --lanes $l --steps 3 --warmup 3 --no-cpu --no-match --no-noise 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['classes_ms_per_step'])"
done; done
bash scripts/gpu_ncu_kernel.sh r02n_describe "k_describe_s" 0 1 16 noise
