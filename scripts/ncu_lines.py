#!/usr/bin/env python
"""Per-source-line executed-instruction / stall-sample profile of one kernel from an ncu report (no GPU needed).
usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-substring> <demangled-substring> <source-file> [top]
Joins `ncu --page source --print-source sass` (executed counts, samples) with `nvdisasm -g` line info by instruction order."""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kern, dkern, srcfile = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l][0]
end = next((i for i, l in enumerate(dis) if i > start and l.startswith('//-----')), len(dis))
cur, per_inst = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l):
        per_inst.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
hdr, inst, first, active = None, [], None, False
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Kernel Name':
        if inst:
            break
        active = dkern in r[1]
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if active and hdr and len(r) == len(hdr):
        if first is None:
            first = r[0]
        elif r[0] == first:
            break
        inst.append(r)
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
assert len(inst) == len(per_inst), (len(inst), len(per_inst))
agg, smp = collections.Counter(), collections.Counter()
for loc, r in zip(per_inst, inst):
    agg[loc] += int(r[iex]); smp[loc] += int(r[ismp])
tot, stot = sum(agg.values()), sum(smp.values())
src = open(srcfile).read().splitlines()
base = srcfile.split('/')[-1]
print(f"{tot} warp instructions executed, {stot} samples, {len(inst)} static")
print(" inst%  smpl%  line  source")
for loc, c in agg.most_common(top):
    f, ln = loc if loc else ('?', 0)
    text = src[ln - 1].strip()[:100] if f == base and 0 < ln <= len(src) else f
    print(f"{100 * c / tot:5.1f} {100 * smp[loc] / max(stot, 1):6.1f} {ln:5d}  {text}")
