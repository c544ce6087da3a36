"""Where a single 1080p frame spends its time: device time of the replayed graph (CUDA events), wall time per C-ABI call,
wall time per Akazer::detectAndCompute call."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab, bindings as B
W, H = 1920, 1080
img = B.u8_to_unit(B.synth_shapes_u8(W, H, seed=1))
for desc in (True, False):
    ctx = ab.Context(W, H, max_batch=1, max_pts=10000)
    dev = torch.from_numpy(img[None]).cuda()
    out = ctx.alloc_results(1, True)
    for _ in range(5):
        ctx.detect_and_compute(dev, desc, out=out)
    ctx.sync()
    s = ctx.torch_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record(s)
    for _ in range(n):
        ctx.detect_and_compute(dev, desc, out=out)
    e1.record(s); e1.synchronize()
    dev_ms = e0.elapsed_time(e1) / n
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.detect_and_compute(dev, desc, out=out)
        ctx.sync()
    wall = (time.perf_counter() - t0) * 1e3 / n
    print(f"describe={desc}: graph replay back to back {dev_ms:.3f} ms/frame (device), call + sync {wall:.3f} ms/frame, {int(out[0][0])} keypoints, launches/frame {ctx.launches // 105}")
    ctx.close()
pitch = (W + 127) // 128 * 128
buf = np.zeros((H, pitch), dtype=np.float32); buf[:, :W] = img
t = torch.from_numpy(buf).cuda()
data = B.DropinData(10000); az = B.DropinAkazer(W, H, pitch)
az.detect_and_compute(t, data); az.time(t, data, iters=5)
print(f"Akazer::detectAndCompute {az.time(t, data, iters=50):.3f} ms/frame; detect only {az.time(t, data, iters=50, desc=False):.3f}")
# per-class device time of one frame (event pairs around every kernel group; no graph, octaves serialised on one stream)
ctx = ab.Context(W, H, max_batch=1, max_pts=10000)
dev = torch.from_numpy(img[None]).cuda()
out = ctx.alloc_results(1, True)
for _ in range(3):
    ctx.detect_and_compute(dev, True, out=out)
ctx.sync(); ctx.profile(True)
ctx.detect_and_compute(dev, True, out=out); ctx.sync()
prof = ctx.profile_read(); ctx.profile(False)
print("classes (ms, launches):", {k: (round(v[0], 3), v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])})
