#!/usr/bin/env python
"""bench_stream.py — BASELINE.json configs[4]: synthetic 4K (3840x2160) stream, 5 octaves x 4 sublevels, end-to-end
detect + describe + match between consecutive frames, the stream split in contiguous ranges over the GPUs with a
one-frame overlap (the boundary frame is recomputed, no exchange: SURVEY 8e).  Run through bench.py:

  python bench.py --config stream [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--stream-frames F]

The stream is one seeded 4K scene panned by (16, 16) pixels per frame (a multiple of 2^4, so every one of the 5 octaves
sees the same sampling lattice): a correct match between frames i-1 and i has a displacement of exactly (16, 16); the
line reports the fraction of accepted matches that satisfy it within 1 px.
A step = the whole stream of this rank: akz_detect_and_compute (raw u8 frames, two lanes) followed on the same stream by
akz_match_pairs (frame f against f - 1 for every f, counts read on the device, reference-compatible cuMatch rule): no host
round trip inside a step.  `value`: frames resident in HBM; `e2e`: pinned host u8 frames in, host keypoints and matches out.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 3840, 2160
NOCT = 5
MAX_PTS = 20000


def stream_frames(rank, world, F):
    import bindings as B
    base = B.synth_shapes_u8(W + 16 * (F * world + 2), H + 16 * (F * world + 2), seed=21, nshapes=600)
    lo = rank * F                                             # this rank's frames are lo .. lo+F-1, plus lo-1 as the overlap frame
    idx = list(range(max(lo - 1, 0), lo + F))
    return np.stack([base[16 * f:16 * f + H, 16 * f:16 * f + W] for f in idx])      # content moves by (-16, -16) per frame


def displacement_accuracy(kp, n, m):
    good = tot = 0
    for f in range(1, len(n)):
        mm = m[f, :n[f]]
        acc = mm[:, 0] >= 0
        q, t = kp[f, :n[f]][acc], kp[f - 1][mm[acc, 0]]
        d = np.stack([t["x"] - q["x"], t["y"] - q["y"]], 1)
        good += int(((np.abs(d[:, 0] - 16) <= 1.0) & (np.abs(d[:, 1] - 16) <= 1.0)).sum())
        tot += int(acc.sum())
    return good, tot


def run(args):
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import akaze_b200 as ab
    import bench as BN
    rank, world, local = BN.dist_setup(args.gpus)
    device = torch.device("cuda", local)
    numa = BN.bind_near_gpu(local, world)
    F = args.stream_frames
    frames = stream_frames(rank, world, F)
    nfr = len(frames)
    host = torch.from_numpy(frames).pin_memory()
    dev = host.to(device)
    chunk = 8
    ctx = ab.Context(W, H, noctaves=NOCT, max_batch=chunk, max_pts=MAX_PTS, device=local, lanes=2)
    stream = ctx.torch_stream()
    res = ctx.alloc_results(nfr, True)
    mres = torch.zeros(nfr, MAX_PTS, 4, dtype=torch.int32, device=device)
    stage = torch.empty_like(dev)
    h_counts = torch.zeros(nfr, dtype=torch.int32).pin_memory()
    h_kpts = torch.zeros((nfr, MAX_PTS, 8), dtype=torch.int32).pin_memory()
    h_m = torch.zeros((nfr, MAX_PTS, 4), dtype=torch.int32).pin_memory()

    def step():
        ctx.detect_and_compute(dev, True, out=res)
        ctx.match_pairs(res[2], res[0], ab.MATCH_COMPAT, out=mres)

    cur = torch.cuda.current_stream(device)

    def step_host():
        # host frames in, host keypoints + matches out, copies inside the timed region.  The copies run on torch's own stream
        # (its pinned-memory allocator tracks that stream), ordered against the library's stream with events.
        stage.copy_(host, non_blocking=True)
        stream.wait_stream(cur)
        ctx.detect_and_compute(stage, True, out=res)
        ctx.match_pairs(res[2], res[0], ab.MATCH_COMPAT, out=mres)
        cur.wait_stream(stream)
        h_counts.copy_(res[0], non_blocking=True)
        cur.synchronize()
        mx = int(h_counts.max())
        if mx > 0:
            h_kpts[:, :mx].copy_(res[1][:, :mx], non_blocking=True)
            h_m[:, :mx].copy_(mres[:, :mx], non_blocking=True)
        cur.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.sync()
    l0 = ctx.launches
    sampler = BN.ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = BN.timed(step, args.steps, stream, world, device)
    launches = ctx.launches - l0
    step_host()
    ms_e2e = BN.timed(step_host, args.steps, cur, world, device)
    clocks = sampler.stop() if rank == 0 else None
    n = res[0].cpu().numpy()
    kp = ab.keypoints_from_words(res[1].cpu().numpy())
    good, tot = displacement_accuracy(kp, n, mres.cpu().numpy())
    assert np.array_equal(n, h_counts.numpy())
    # roofline of the dominant kernel class (same definition as the frames config, 4K level sizes)
    ctx.profile(True)
    for f0 in range(0, nfr, chunk):
        ctx.detect_and_compute(dev[f0:f0 + chunk], True, out=tuple(r[f0:f0 + chunk] for r in res))
    ctx.match_pairs(res[2], res[0], ab.MATCH_COMPAT, out=mres)
    prof = ctx.profile_read()
    ctx.profile(False)
    px = BN.level_pixels(W, H, NOCT)
    lv = sum(px)
    alg_cls = {"prep": 20 * lv - 4 * px[0] + 4 * sum(px[i] for i in range(4, len(px), 4)), "fed": 12 * (lv - px[0]), "base": 9 * px[0]}
    top = max((k for k in prof if k in alg_cls), key=lambda k: prof[k][0])
    top_ms, top_launches = prof[top]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = alg_cls[top] * nfr
    achieved = alg / (top_ms * 1e-3) / 1e9
    step_alg = (W * H + 20 * lv) * nfr
    roof = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "kernel_ms_per_step": round(top_ms, 3), "kernel_launches_per_step": int(top_launches), "algorithmic_bytes_per_launch": alg / max(1, top_launches),
            "classes_ms_per_step": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "whole_step": {"algorithmic_gbs": round(step_alg / (ms / args.steps * 1e-3) / 1e9, 1)}}
    nframes_out = (nfr - 1) * world
    line = {"metric": "4K stream detect+describe+match frames/sec", "value": round(nframes_out * args.steps / (ms * 1e-3), 2), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: synthetic 3840x2160 u8 stream, 5 octaves x 4 sublevels, detect+describe+match of consecutive frames "
                                   "(cuMatch rule), stream split over the GPUs with a one-frame overlap, no collective",
                       "frames_per_gpu": F, "chunk": chunk, "lanes": 2, "max_pts": MAX_PTS, "levels": ctx.num_levels,
                       "keypoints_per_frame_mean": round(float(n.mean()), 1), "binding": numa,
                       "l2_policy": f"inputs larger than L2: {nfr} frames x {W * H / 1e6:.1f} MB in, {chunk} x 810 MB of planes per chunk"},
            "e2e": {"value": round(nframes_out * args.steps / (ms_e2e * 1e-3), 2), "unit": "frames/s", "h2d_bytes_per_step": int(host.numel()),
                    "d2h_bytes_per_step": int(nfr * 4 + n.sum() * (32 + 16)), "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "matches_accepted": tot, "matches_with_the_true_displacement": round(good / max(tot, 1), 4)}
    torch.cuda.synchronize()
    del stream, cur
    ctx.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_stream_baseline(frames[:3])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_stream_baseline(sample):
    """OpenCV cv::AKAZE (5 octaves) + BFMatcher on the first frames of the stream, all host cores."""
    cores = os.cpu_count() or 1
    try:
        import cv2
        cv2.setNumThreads(cores)
        ak = cv2.AKAZE_create(nOctaves=NOCT)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        t0 = time.perf_counter()
        prev = None
        nk = 0
        for f in sample:
            k, d = ak.detectAndCompute(f, None)
            nk += len(k)
            if prev is not None and d is not None and len(d):
                bf.match(d, prev)
            prev = d
        dt = time.perf_counter() - t0
        return {"value": round((len(sample) - 1) / dt, 3), "unit": "frames/s", "cores": cores, "kind": "reference",
                "sample": f"cv2 {cv2.__version__} AKAZE_create(nOctaves=5) + BFMatcher on {len(sample)} frames of the stream, {nk / len(sample):.0f} keypoints/frame"}
    except Exception as e:
        return {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "cv2 unavailable: " + repr(e)[:100]}


def run_reference(args):
    """The unmodified reference on the same stream: Akazer::detectAndCompute per frame (5 octaves) + cuMatch per consecutive pair.
    gHammingMatch never returns on sm_100a when the train count is not a multiple of 16 (App. B-15): the train set of every
    pair is cut to the multiple of 16 below its count, which is noted in the line."""
    import ctypes as C
    import torch
    import bindings as B
    import bench as BN
    rank, world, local = BN.dist_setup(args.gpus)
    device = torch.device("cuda", local)
    if not B.have_ref():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_akaze.so not built (needs /root/reference at build time)"}))
        return
    F = min(args.stream_frames, 8)
    frames = stream_frames(rank, world, F)
    nfr = len(frames)
    pitch = (W + 127) // 128 * 128
    host = torch.zeros((nfr, H, pitch), dtype=torch.float32)
    host.numpy()[:, :, :W] = frames.astype(np.float32) * np.float32(1.0 / 255.0)
    dev = host.to(device)
    ref = B.RefAkazer(W, H, pitch, noctaves=NOCT)
    L = ref.L
    pts = [torch.zeros(MAX_PTS * 104, dtype=torch.uint8, device=device) for _ in range(2)]
    hq = torch.zeros(MAX_PTS * 104, dtype=torch.uint8).pin_memory()
    counts = np.zeros(nfr, dtype=np.int64)

    def step():
        for f in range(nfr):
            counts[f] = L.ref_akazer_detectAndCompute(ref.hnd, C.c_void_p(dev[f].data_ptr()), W, H, pitch, 1, C.c_void_p(pts[f & 1].data_ptr()), None, MAX_PTS)
            if f > 0 and counts[f] > 0 and counts[f - 1] >= 16:
                L.ref_cuMatch(C.c_void_p(pts[f & 1].data_ptr()), C.c_void_p(hq.data_ptr()), int(counts[f]),
                              C.c_void_p(pts[(f - 1) & 1].data_ptr()), int(counts[f - 1]) // 16 * 16)

    stream = torch.cuda.current_stream()
    step()
    steps = max(1, args.steps // 2)
    ms = BN.timed(step, steps, stream, world, device)
    v = (nfr - 1) * world * steps / (ms * 1e-3)
    line = {"impl": "reference", "metric": "4K stream detect+describe+match frames/sec", "value": round(v, 2), "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": 1, "ms_per_step": round(ms / steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: synthetic 3840x2160 stream, 5 octaves x 4 sublevels, detect+describe+match of consecutive frames",
                       "frames_per_gpu": F, "max_pts": MAX_PTS, "keypoints_per_frame_mean": round(float(counts.mean()), 1),
                       "how": "unmodified reference compiled for sm_100a: Akazer::detectAndCompute per frame + cuMatch per pair (train count cut to a "
                              "multiple of 16: gHammingMatch deadlocks otherwise); float frames resident on the device"},
            "e2e": {"value": round(v, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(counts.sum() * 16)},
            "cpu_baseline": {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                             "sample": "not a CPU run: the reference implementation of this path is CUDA (cv::AKAZE numbers are in the main arm's cpu_baseline)"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ref.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    import bench
    sys.argv = [sys.argv[0], "--config", "stream"] + sys.argv[1:]
    bench.main()
