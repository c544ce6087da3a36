"""One FED cycle configuration, a few launches (for ncu): python scripts/fed_one.py W H NF N"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab
w, h, nf, n = [int(x) for x in sys.argv[1:5]]
ctx = ab.Context(0, 0, fused=1, max_batch=nf)
L = torch.rand(nf, h, w, device="cuda"); G = torch.rand(nf, h, w, device="cuda")
dst, tmp = torch.zeros_like(L), torch.zeros_like(L)
tau = (np.random.default_rng(n).random(n) * 0.2 + 0.01).astype(np.float32)
for _ in range(4):
    ctx.fed_cycle(L, G, dst, tmp, w, tau)
ctx.sync()
print("ok")
