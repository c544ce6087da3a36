"""How many pixels pass the detector threshold / are 3x3 maxima per level of one synthetic 1080p frame (shapes and noise)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab, bench as BN
for content in ("shapes", "noise"):
    fr = BN.make_frames(1, content)
    dev = (torch.from_numpy(fr).cuda().float() * (1.0 / 255.0)).contiguous()
    ctx = ab.Context(BN.W, BN.H, max_batch=1, max_pts=32768)
    ctx.build_scale_space(dev); ctx.sync()
    tot_thr = tot_max = 0
    for l in range(ctx.num_levels):
        d = ctx.plane(l, ab.PLANE_DET)
        d = d.cpu().numpy() if hasattr(d, "cpu") else np.asarray(d)
        if d is None: break
        a = d[1:-1, 1:-1]
        thr = a > 0.001
        mx = thr.copy()
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dy or dx: mx &= a > d[1 + dy:d.shape[0] - 1 + dy, 1 + dx:d.shape[1] - 1 + dx]
        tot_thr += int(thr.sum()); tot_max += int(mx.sum())
        print(content, "level", l, d.shape, "above threshold", int(thr.sum()), f"({thr.mean():.3f})", "3x3 maxima", int(mx.sum()), f"({mx.mean():.4f})")
    print(content, "total above", tot_thr, "maxima", tot_max)
    ctx.close()
