import sys
sys.path.insert(0, "cuda-akaze_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
import akaze_b200 as ab
import bindings as B
L = ab.lib(); ctx = ab.Context(0, 0)
for nq, nt in ((10000, 10000), (10000, 100000), (10000, 1000000)):
    q = torch.from_numpy(B.random_descriptors(nq, 0)).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    t = torch.randint(0, 256, (nt, 64), dtype=torch.uint8, device="cuda", generator=g); t[:, 61:] = 0; t[:, 60] &= 0x3F
    res = torch.zeros(nq, 4, dtype=torch.int32, device="cuda")
    for mode, name in ((ab.MATCH_KNN2, "knn2"), (ab.MATCH_COMPAT, "compat")):
        L.akz_set_match_kernel(3)
        for _ in range(3): ctx.match(q, t, mode, out=res)
        ctx.sync(); s = ctx.torch_stream()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): ctx.match(q, t, mode, out=res)
        e1.record(s); e1.synchronize()
        print(f"{nq}x{nt} {name}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
