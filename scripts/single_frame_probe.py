"""Latency of ONE 1080p frame through a max_batch = 1 context (the shape of Akazer::detectAndCompute calls in main.cpp)."""
import sys, time
sys.path.insert(0, "cuda-akaze_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import akaze_b200 as ab, bench as BN
frames8 = BN.make_frames(4, "shapes")
dev = torch.from_numpy(frames8.astype(np.float32) * np.float32(1 / 255.)).cuda()
ctx = ab.Context(BN.W, BN.H, max_batch=1, max_pts=10000)
res = ctx.alloc_results(1, True)
for i in range(5): ctx.detect_and_compute(dev[i % 4:i % 4 + 1], True, out=res)
ctx.sync()
s = ctx.torch_stream()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
l0 = ctx.launches
t0 = time.perf_counter(); e0.record(s)
N = 50
for i in range(N): ctx.detect_and_compute(dev[i % 4:i % 4 + 1], True, out=res)
t1 = time.perf_counter()
e1.record(s); e1.synchronize(); t2 = time.perf_counter()
print(f"single frame: {e0.elapsed_time(e1) / N:.3f} ms on the device stream, host enqueue {1e3 * (t1 - t0) / N:.3f} ms/frame, wall {1e3 * (t2 - t0) / N:.3f} ms/frame, {(ctx.launches - l0) / N:.0f} launches/frame")
