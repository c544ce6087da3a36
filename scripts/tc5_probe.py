"""GPU probe of the tcgen05 matcher (kernel 3) against the LOP3/POPC kernel (1) and the mma.sync kernel (2)."""
import sys, time
sys.path.insert(0, "cuda-akaze_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import akaze_b200 as ab
import bindings as B

L = ab.lib()
ctx = ab.Context(0, 0)
ok_all = True
for nq, nt in ((128, 128), (100, 300), (1000, 1500), (300, 129), (2049, 4097), (5, 77), (4000, 9000)):
    q = B.random_descriptors(nq, nq + nt); t = B.random_descriptors(nt, nq + nt + 1)
    t[: min(nq, nt) // 3] = q[: min(nq, nt) // 3]
    if nt > 40: t[20] = t[4]; t[21] = t[5]
    qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    for mode in (ab.MATCH_COMPAT, ab.MATCH_KNN2):
        out = {}
        for kern in (1, 3, 4, 5):
            L.akz_set_match_kernel(kern)
            r = ctx.match(qt, tt, mode); ctx.sync()
            out[kern] = r.cpu().numpy()
        same = all(np.array_equal(out[1], out[k]) for k in (3, 4, 5))
        ok_all &= same
        bad = np.argwhere(np.logical_or.reduce([(out[1] != out[k]).any(axis=1) for k in (3, 4, 5)]))[:, 0]
        print(f"{nq}x{nt} mode {mode}: equal={same} mismatching rows={len(bad)}", flush=True)
        if len(bad):
            for i in bad[:4]: print("   row", i, out[1][i], out[3][i], out[4][i], out[5][i])
print("ALL EQUAL" if ok_all else "MISMATCH", flush=True)

for nq, nt in ((10000, 10000), (10000, 100000)):
    q = B.random_descriptors(nq, 0); t = B.random_descriptors(nt, 1)
    qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    for kern in (1, 2, 3, 4, 5):
        L.akz_set_match_kernel(kern)
        for _ in range(3): ctx.match(qt, tt, ab.MATCH_KNN2)
        ctx.sync()
        s = ctx.torch_stream()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): ctx.match(qt, tt, ab.MATCH_KNN2)
        e1.record(s); e1.synchronize()
        print(f"{nq}x{nt} kernel {kern}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
L.akz_set_match_kernel(0)
q = B.random_descriptors(10000, 0); t = B.random_descriptors(10000, 1)
qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
for kern in (3, 4, 5):
    L.akz_set_match_kernel(kern)
    for _ in range(3): ctx.match(qt, tt, ab.MATCH_COMPAT)
    ctx.sync()
    s = ctx.torch_stream()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): ctx.match(qt, tt, ab.MATCH_COMPAT)
    e1.record(s); e1.synchronize()
    print(f"compat 10000x10000 kernel {kern}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
L.akz_set_match_kernel(0)
