#!/bin/bash
# short bench line: classes of the default and of the noise workload
python bench.py --frames 128 --steps 3 --warmup 3 --no-cpu --no-match --noise-frames 32 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], d['roofline']['classes_ms_per_step'])
print('noise', d['noise']['value'], d['noise']['classes_ms_per_step'])
"
