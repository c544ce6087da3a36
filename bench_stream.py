#!/usr/bin/env python
"""bench_stream.py — BASELINE.json configs[4]: synthetic 4K (3840x2160) stream, 5 octaves x 4 sublevels, end-to-end
detect + describe + match between consecutive frames, the stream split in contiguous ranges over the GPUs with a
one-frame overlap (the boundary frame is recomputed, no exchange: SURVEY 8e).

  python bench_stream.py [--frames F] [--chunk C] [--steps K]          (N > 1: launch with torch.distributed.run)

The stream is one seeded 4K scene panned by (16, 16) pixels per frame (a multiple of 2^4, so every one of the 5 octaves
sees the same sampling lattice): a correct match between frames i-1 and i has a displacement of exactly (16, 16); the
script reports the fraction of accepted matches that satisfy it within 1 px.
Frames are raw u8 resident in HBM when the timed region starts; matching uses the reference-compatible mode (cuMatch).
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 3840, 2160


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32, help="frames of the stream per GPU (plus the overlap frame)")
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--max-pts", type=int, default=20000)
    args = ap.parse_args()
    import torch
    import akaze_b200 as ab
    import bindings as B
    import bench as BN
    rank, world, local = BN.dist_setup(world := int(os.environ.get("WORLD_SIZE", "1")))
    device = torch.device("cuda", local)
    F = args.frames
    base = B.synth_shapes_u8(W + 16 * (F * world + 2), H + 16 * (F * world + 2), seed=21, nshapes=600)
    lo = rank * F                                             # this rank's frames are lo .. lo+F-1, plus lo-1 as the overlap frame
    idx = list(range(max(lo - 1, 0), lo + F))
    frames = np.stack([base[16 * f:16 * f + H, 16 * f:16 * f + W] for f in idx])      # content moves by (-16, -16) per frame
    dev = torch.from_numpy(frames).to(device)
    nfr = len(idx)
    ctx = ab.Context(W, H, noctaves=5, max_batch=args.chunk, max_pts=args.max_pts, device=local, lanes=2)
    mctx = ab.Context(0, 0, device=local)
    res = ctx.alloc_results(nfr, True)
    mres = torch.zeros(nfr, args.max_pts, 4, dtype=torch.int32, device=device)

    def step():
        counts, kpts, desc = ctx.detect_and_compute(dev, True, out=res)
        ctx.sync()
        n = counts.cpu().numpy()
        for f in range(1, nfr):
            if n[f] > 0 and n[f - 1] >= 16:
                mctx.match(desc[f, :n[f]], desc[f - 1, :n[f - 1]], ab.MATCH_COMPAT, out=mres[f, :n[f]])
        mctx.sync()
        return n

    for _ in range(2):
        n = step()
    BN.barrier(world)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        n = step()
    e1.record()
    e1.synchronize()
    BN.barrier(world)
    ms = BN.barrier_max(e0.elapsed_time(e1), world, device)
    # check the geometry of the accepted matches
    kp = ab.keypoints_from_words(res[1].cpu().numpy())
    m = mres.cpu().numpy()
    good = tot = 0
    for f in range(1, nfr):
        mm = m[f, :n[f]]
        acc = mm[:, 0] >= 0
        q, t = kp[f, :n[f]][acc], kp[f - 1][mm[acc, 0]]
        d = np.stack([t["x"] - q["x"], t["y"] - q["y"]], 1)
        good += int(((np.abs(d[:, 0] - 16) <= 1.0) & (np.abs(d[:, 1] - 16) <= 1.0)).sum())
        tot += int(acc.sum())
    line = {"metric": "4K stream detect+describe+match frames/sec", "value": round((nfr - 1) * world * args.steps / (ms * 1e-3), 2),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "dtype": "u8 in, f32 arithmetic", "data": "synthetic",
            "config": {"workload": "configs[4]: synthetic 3840x2160 stream, 5 octaves x 4 sublevels, detect+describe+match consecutive frames",
                       "frames_per_gpu": F, "chunk": args.chunk, "keypoints_per_frame_mean": round(float(n.mean()), 1),
                       "levels": ctx.num_levels, "launches_per_step": None},
            "matches_accepted": tot, "matches_with_the_true_displacement": round(good / max(tot, 1), 4)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close(); mctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
