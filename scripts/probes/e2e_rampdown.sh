for rd in 0 1; do
if [ $rd = 1 ]; then export AKZ_RAMP_DOWN=1; fi
python bench.py --steps 4 --warmup 3 --no-cpu --no-match --no-noise 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rd=$rd value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'e2e_u8', d['e2e_u8']['value'])"
done
