#!/usr/bin/env python
"""10k x 10k brute-force matching a few times (the command ncu wraps for the matcher captures).
usage: match_probe.py [kernel [nq nt]]   kernel: 0 = by size, 1 = LOP3/POPC, 2 = mma.sync, 3 = tcgen05"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import akaze_b200 as ab
import bindings as B
ab.lib().akz_set_match_kernel(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ctx = ab.Context(0, 0)
NQ = int(sys.argv[2]) if len(sys.argv) > 3 else 10000
NT = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
q = torch.from_numpy(B.random_descriptors(NQ, 0)).cuda()
t = torch.from_numpy(B.random_descriptors(NT, 1)).cuda()
MODES = (ab.MATCH_KNN2,) * 3 if len(sys.argv) > 4 and sys.argv[4] == "knn2" else (ab.MATCH_KNN2, ab.MATCH_COMPAT, ab.MATCH_KNN2)
for mode in MODES:
    r = ctx.match(q, t, mode)
ctx.sync()
print("ok", int(r[:, 1].min()))
