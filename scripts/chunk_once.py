"""One chunk of synthetic 1080p frames through akz_detect_and_compute, a few times (for ncu launch lists and captures):
   python scripts/chunk_once.py [frames=32] [iters=3] [content=shapes]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab, bench as BN
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
content = sys.argv[3] if len(sys.argv) > 3 else "shapes"
frames = BN.make_frames(nf, content)
dev = (torch.from_numpy(frames).cuda().float() * (1.0 / 255.0)).contiguous()
ctx = ab.Context(BN.W, BN.H, max_batch=nf, max_pts=32768 if content == "noise" else 10000, lanes=1)
res = ctx.alloc_results(nf, True)
for _ in range(iters):
    ctx.detect_and_compute(dev, True, out=res)
ctx.sync()
print("ok", int(res[0].sum()))
