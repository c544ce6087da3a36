#!/usr/bin/env python
"""bench_fast.py — throughput of the integer ("fast") pipeline, Akazer::fastDetectAndCompute, on the
configs[2] workload (synthetic 1920x1080 u8 frames resident in HBM): this repo's akz_fast_detect_and_compute against the
unmodified reference's fastDetectAndCompute (oracle/_ref/libref_akaze.so).  One JSON line per arm."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import akaze_b200 as ab
    import bindings as B
    import bench as BN
    F, W, H = args.frames, BN.W, BN.H
    frames8 = BN.make_frames(F, "shapes")
    dev = torch.from_numpy(frames8).cuda()
    ctx = ab.Context(W, H, max_batch=args.chunk, max_pts=20000, lanes=2)
    res = ctx.alloc_results(F, True)
    stream = ctx.torch_stream()
    step = lambda: ctx.fast_detect_and_compute(dev, True, out=res)
    for _ in range(2):
        step()
    ctx.sync()
    ms = BN.timed(step, args.steps, stream, 1, dev.device)
    n = res[0].cpu().numpy()
    ctx.profile(True)
    for f0 in range(0, F, args.chunk):
        ctx.fast_detect_and_compute(dev[f0:f0 + args.chunk], True, out=tuple(r[f0:f0 + args.chunk] for r in res))
    prof = ctx.profile_read()
    ctx.profile(False)
    classes = {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[0] > 0}
    print(json.dumps({"metric": "1080p integer-pipeline detect+describe images/sec", "impl": "ours", "classes_ms_per_step": classes, "value": round(F * args.steps / (ms * 1e-3), 2),
                      "unit": "images/s", "ms_per_step": round(ms / args.steps, 3), "frames": F, "keypoints_per_frame_mean": round(float(n.mean()), 1),
                      "note": "fused level kernel k_prep2<.., INT> and temporally blocked k_fed3<int> of the float pipeline instantiated for int32 planes, two lanes, batched, no host round trips"}), flush=True)
    ctx.close()
    if B.have_ref():
        pitch = (W + 127) // 128 * 128
        buf = np.zeros((F, H, pitch), dtype=np.uint8)
        buf[:, :, :W] = frames8
        rdev = torch.from_numpy(buf).cuda()
        ref = B.RefAkazer(W, H, pitch)
        L = ref.L
        pts = torch.zeros(20000 * 104, dtype=torch.uint8, device="cuda")
        counts = np.zeros(F)

        def rstep():
            for f in range(F):
                counts[f] = L.ref_akazer_fastDetectAndCompute(ref.hnd, C.c_void_p(rdev[f].data_ptr()), W, H, pitch, 1, C.c_void_p(pts.data_ptr()), None, 20000)
        rstep()
        ms = BN.timed(rstep, max(1, args.steps - 1), torch.cuda.current_stream(), 1, rdev.device)
        print(json.dumps({"metric": "1080p integer-pipeline detect+describe images/sec", "impl": "reference", "value": round(F * max(1, args.steps - 1) / (ms * 1e-3), 2),
                          "unit": "images/s", "frames": F, "keypoints_per_frame_mean": round(float(counts.mean()), 1)}), flush=True)
        ref.close()


if __name__ == "__main__":
    main()
