#!/usr/bin/env python
"""10k x 10k brute-force matching a few times (the command ncu wraps for the matcher capture)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import akaze_b200 as ab
import bindings as B
ctx = ab.Context(0, 0)
q = torch.from_numpy(B.random_descriptors(10000, 0)).cuda()
t = torch.from_numpy(B.random_descriptors(10000, 1)).cuda()
for mode in (ab.MATCH_KNN2, ab.MATCH_COMPAT, ab.MATCH_KNN2):
    r = ctx.match(q, t, mode)
ctx.sync()
print("ok", int(r[:, 1].min()))
