import sys, os, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab, bindings as B

def desc(name):
    img = B.u8_to_unit(B.read_pgm(os.path.join(B.REF_DATA, name)))
    h, w = img.shape
    ctx = ab.Context(w, h, max_batch=1, max_pts=10000)
    c, k, d = ctx.detect_and_compute(torch.from_numpy(img[None]).cuda())
    ctx.sync(); n = int(c[0].cpu())
    kk = ab.keypoints_from_words(k[0, :n].cpu().numpy()); dd = d[0, :n].cpu().numpy(); ctx.close()
    return kk, dd
k1, d1 = desc("left.pgm"); k2, d2 = desc("right.pgm")
n2 = len(d2) // 16 * 16; d2 = d2[:n2]; k2 = k2[:n2]
print(len(d1), len(d2))
o = B.oracle_match(d1, d2, "compat")
ctx = ab.Context(0, 0)
for kern in (1, 2, 3):
    ab.lib().akz_set_match_kernel(kern)
    r = ctx.match(torch.from_numpy(d1).cuda(), torch.from_numpy(d2).cuda(), ab.MATCH_COMPAT).cpu().numpy()
    print("kernel", kern, "vs oracle: idx mismatches", int((r[:, 0] != o[:, 0]).sum()), "dist mismatches", int((r[:, 1] != o[:, 1]).sum()))
ab.lib().akz_set_match_kernel(0)
pq = np.zeros(len(d1), dtype=B.REF_POINT); pt = np.zeros(n2, dtype=B.REF_POINT)
pq["features"], pt["features"] = d1[:, :61], d2[:, :61]
pt["x"], pt["y"] = k2["x"], k2["y"]
dq = torch.from_numpy(pq.view(np.uint8).reshape(-1)).cuda(); dt = torch.from_numpy(pt.view(np.uint8).reshape(-1)).cuda()
torch.cuda.synchronize()
B.ref().ref_hMatch(C.c_void_p(dq.data_ptr()), len(d1), C.c_void_p(dt.data_ptr()), n2)
torch.cuda.synchronize()
rq = dq.cpu().numpy().view(B.REF_POINT)
a = o[:, 0]; b = rq["match"]
print("oracle vs ref: agree", (a == b).mean(), "oracle>=0,ref<0:", int(((a >= 0) & (b < 0)).sum()), "oracle<0,ref>=0:", int(((a < 0) & (b >= 0)).sum()),
      "both>=0 differ:", int(((a >= 0) & (b >= 0) & (a != b)).sum()))
bad = np.where((a >= 0) & (b < 0) & (o[:, 1] < 72))[0]
print("oracle accepted with dist<72 but ref rejected:", len(bad))
for i in bad[:10]:
    # brute force distances
    x = np.unpackbits(d1[i][None] ^ d2, axis=1).sum(1)
    srt = np.argsort(x, kind="stable")[:4]
    print(i, "oracle", o[i], "best", [(int(j), int(x[j]), int(j) % 16) for j in srt])
both = (a >= 0) & (b >= 0)
off = rq["distance"][both] - o[both, 1]
print("distance offsets (ref - oracle):", np.unique(off, return_counts=True))
