"""Timing + bit-exactness probe of the FED cycle kernels at pyramid sizes (AKZ_FED_STREAM=0/1, AKZ_FED_BAND)."""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cuda-akaze_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import akaze_b200 as ab

def run(w, h, nf, n, check):
    ctx = ab.Context(0, 0, fused=1, max_batch=nf)
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    L = torch.rand(nf, h, w, device="cuda", generator=g); G = torch.rand(nf, h, w, device="cuda", generator=g)
    tau = (np.random.default_rng(n).random(n) * 0.2 + 0.01).astype(np.float32)
    dst, tmp = torch.zeros_like(L), torch.zeros_like(L)
    for _ in range(2):
        ctx.fed_cycle(L, G, dst, tmp, w, tau)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = ctx.torch_stream()
    reps = 5
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(reps):
            ctx.fed_cycle(L, G, dst, tmp, w, tau)
        e1.record(s)
    ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    ok = ""
    if check:
        c0 = ab.Context(0, 0, fused=0, max_batch=nf)
        d2, t2 = torch.zeros_like(L), torch.zeros_like(L)
        c0.fed_cycle(L, G, d2, t2, w, tau); c0.sync()
        ok = "exact" if torch.equal(dst.view(torch.int32), d2.view(torch.int32)) else "MISMATCH %d" % int((dst.view(torch.int32) != d2.view(torch.int32)).sum())
        c0.close()
    px = w * h * nf * n
    print(f"{w}x{h} x{nf} n={n}: {ms:.3f} ms  {px / ms / 1e6:.1f} G step-px/s  {12.0 * w * h * nf / ms / 1e6:.0f} GB/s alg  {ok}", flush=True)
    ctx.close()
    return ms

if __name__ == "__main__":
    check = "--check" in sys.argv
    tot = 0.0
    for (w, h, steps) in ((1920, 1080, (3, 3, 4)), (960, 540, (4, 5, 6, 7)), (480, 270, (8, 10, 12, 14)), (240, 135, (17, 20, 24, 29))):
        for n in steps:
            tot += run(w, h, 32, n, check)
    print(f"sum over the 15 cycles of a 32-frame chunk: {tot:.3f} ms -> {tot * 8:.2f} ms per 256 frames")
