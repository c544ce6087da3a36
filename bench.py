#!/usr/bin/env python
"""bench.py — 1080p AKAZE detect + describe throughput (BASELINE.json metric 1) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--content shapes|noise]

A "step" is one pass of the hot path (akz_detect_and_compute: nonlinear scale space, Hessian detector, NMS,
refinement, orientation, M-LDB) over one batch of F synthetic 1920x1080 frames per GPU (BASELINE.json configs[2]:
batch of 256 frames, frame-sharded: weak scaling, no data-path collective).  One JSON line is printed by rank 0:

  value         frames/s with the batch already resident in HBM (device pointers in, device results out)
  e2e           frames/s through akz_detect_and_compute_host: pinned HOST frames in, HOST keypoints/descriptors
                out, H2D and D2H copies inside the timed region
  roofline      the dominant kernel class of the step, timed with CUDA events on the library's stream
                (akz_profile_*) in one extra pass over the same batch; achieved = algorithmic bytes / time
  cpu_baseline  OpenCV cv::AKAZE (the reference's own CPU arm, main.cpp:344-399) and the C oracle port, on a
                bounded sample of the same frames on the box's host cores (rank 0, N=1 only)

  noise         the same step on SURVEY 8(d)'s keypoint-heavy input (blurred uniform noise, ~21 k keypoints per frame, max_pts
                32768) with its own per-class times: the default "shapes" frames carry ~1.8 k keypoints and hide the keypoint stages
  single_frame  ms per frame of the synchronous drop-in entry point akaze::Akazer::detectAndCompute (include/akaze.h) on one
                resident 1080p frame -- the reference's own timed loop, main.cpp:199-205
  match         BASELINE metric 2: 10k x 10k on one GPU; 10k x 1M with the train set sharded over the ranks through
                akz_match_sharded (one ncclAllGather inside the library), asserted equal to the unsharded result

--config stream runs BASELINE configs[4] instead (synthetic 3840x2160 stream, 5 octaves x 4 sublevels, detect + describe +
match of consecutive frames, the stream split over the GPUs with a one-frame overlap).

--impl reference runs the UNMODIFIED reference CUDA library (oracle/_ref/libref_akaze.so, built from
/root/reference by oracle/Makefile for sm_100a) through its own Akazer::detectAndCompute, frame by frame as
main.cpp:199-205 does, on the same frames.  The reference has no CPU implementation of its own.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 1920, 1080
L2_BYTES = 126 * 1024 * 1024


def level_pixels(w, h, noct=4, S=4):
    px = []
    for o in range(noct):
        ww, hh = w >> o, h >> o
        if o and (ww < 80 or hh < 80):
            break
        px += [ww * hh] * S
    return px


def make_frames(nframes, content, seed0=0):
    """F distinct u8 frames: 8 seeded base images (tests/bindings.py generators), each rolled by a
    different offset so no two frames of the batch are equal."""
    import bindings as B
    nbase = min(8, nframes)
    gen = B.synth_noise_u8 if content == "noise" else B.synth_shapes_u8
    base = [gen(W, H, seed=seed0 + s) for s in range(nbase)]
    out = np.empty((nframes, H, W), dtype=np.uint8)
    for f in range(nframes):
        b = base[f % nbase]
        k = f // nbase
        out[f] = np.roll(b, (37 * k, 53 * k), axis=(0, 1)) if k else b
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smmax.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(ngpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def bind_near_gpu(local, world):
    """Pin this process (and the pinned buffers it allocates afterwards: first touch) to the CPUs of the GPU's NUMA node,
    and when several ranks share a node give each its own slice of those CPUs.  Returns a short description."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = []
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0)) or sorted(os.sched_getaffinity(0))
        if world > 1 and len(allowed) >= world:
            per = len(allowed) // world
            allowed = allowed[local * per:(local + 1) * per]
        os.sched_setaffinity(0, allowed)
        return {"pci": bdf, "numa_node": node, "cpus": f"{allowed[0]}-{allowed[-1]}" if allowed else "", "ncpus": len(allowed)}
    except Exception as e:
        return {"error": repr(e)[:100]}


def barrier_max(ms, world, device):
    import torch
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, stream, world, device):
    """barrier + sync | K steps between two CUDA events on `stream` | barrier + sync; max over ranks (ms)."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    e1.synchronize()
    barrier(world)
    return barrier_max(e0.elapsed_time(e1), world, device)


# ---- algorithmic bytes per kernel class, per frame (DESIGN.md "roofline") ---------------------------------------
def algorithmic_bytes(cls, nkp):
    px = level_pixels(W, H)
    lv = sum(px)
    if cls == "fed":       # per FED cycle: read Lt_prev and g, write Lt  (12 B/px), cycles = every level but (0,0)
        return 12 * (lv - px[0])
    if cls == "prep":      # read Lt_prev, write g, Lx, Ly, det (20 B/px); octave transitions also write Lt (+4)
        return 20 * lv - 4 * px[0] + 4 * sum(px[i] for i in range(4, len(px), 4))
    if cls == "base":      # read the frame twice-in-one (4 B/px), write smooth(sigma 1) and Lt(0,0)
        return 12 * px[0]
    if cls == "contrast":
        return 4 * px[0]
    if cls == "describe":  # 441 samples x 3 planes x 4 B + 64 B out per keypoint
        return nkp * (441 * 12 + 64)
    if cls == "orient":
        return nkp * (109 * 8 + 32)
    if cls == "extrema":
        return 4 * lv
    if cls == "nms":
        return 8 * px[0]
    return 0


def match_metric(ab, device, rank, world, clock_mhz):
    """BASELINE metric 2: brute-force Hamming matching, 10k x 10k on one GPU and 10k x 1M with the TRAIN set sharded over
    the ranks (akz_match(finalize=0) per shard -> one all_gather of nq x 16 B per rank over NCCL -> akz_match_merge)."""
    import torch
    import bindings as B
    from akaze_b200 import distributed as D
    ctx = ab.Context(0, 0, device=device.index)
    stream = ctx.torch_stream()
    out = {}
    q = torch.from_numpy(B.random_descriptors(10000, 0)).to(device)
    t = torch.from_numpy(B.random_descriptors(10000, 1)).to(device)
    res = torch.zeros(10000, 4, dtype=torch.int32, device=device)

    def time_it(fn, iters):
        for _ in range(3):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / iters

    for mode, name in ((ab.MATCH_KNN2, "knn2"), (ab.MATCH_COMPAT, "compat")):
        out[f"{name}_10kx10k_ms"] = round(time_it(lambda: ctx.match(q, t, mode, out=res), 20), 4)
    # SURVEY 8d counts 16 XOR + 16 POPC per pair (the plain formulation, POPC bound at 16 POPC/clk/SM).  Large problems run on
    # the tcgen05 kernel (u8 GEMM on bit-expanded descriptors, 512 MACs per pair), so two fractions are reported: pairs/s
    # against the plain formulation's POPC bound (can exceed 1: the tensor cores are not bound by it) and MACs/s against the
    # dense int8 tensor rate of the tcgen05 floor (128 x 128 x 32 MACs per 64 clocks per SM, B300_MICROARCH.md "tcgen05 floor").
    pairs = 10000 * 10000 / (out["knn2_10kx10k_ms"] * 1e-3)
    clk = (clock_mhz or 1965.0) * 1e6
    peak_pairs = 148 * 16 * clk / 16
    peak_macs = 148 * 8192 * clk
    out["kernel"] = "k_match_tc5 (tcgen05.mma kind::i8, TMEM accumulators) + k_match_merge"
    out["pairs_per_s"] = float(f"{pairs:.4g}")
    out["frac_of_plain_popc_bound"] = round(pairs / peak_pairs, 3)
    out["plain_popc_bound"] = "148 SMs x 16 POPC.32/clk x SM clock / 16 POPC per pair"
    out["frac_of_i8_tensor_peak"] = round(pairs * 512 / peak_macs, 3)
    out["i8_tensor_peak"] = "148 SMs x 8192 MAC/clk x SM clock (tcgen05 kind::i8, M = 128)"
    # 10k x 1M, train sharded over the ranks (config 4)
    nt_total = 1_000_000
    lo, hi = D.shard_bounds(nt_total, world, rank)
    g = torch.Generator(device=device)
    g.manual_seed(1234 + rank)
    tl = torch.randint(0, 256, (hi - lo, 64), dtype=torch.uint8, device=device, generator=g)
    tl[:, 61:] = 0
    tl[:, 60] &= 0x3F
    if world > 1:
        import torch.distributed as dist
        D.comm_init_from_group(ctx)                       # the library's own communicator: id from rank 0, broadcast by torch
        fn = lambda: ctx.match_sharded(q, tl, lo, ab.MATCH_KNN2, out=res)
        for _ in range(3):
            fn()
        ctx.sync()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            fn()
        e1.record(stream)
        e1.synchronize()
        ms = barrier_max(e0.elapsed_time(e1) / 10, world, device)
        # the sharded result must equal the unsharded one: gather the whole train set once (outside the timed region) and match
        # it on this rank alone; also in the reference-compatible mode
        sizes = [D.shard_bounds(nt_total, world, r)[1] - D.shard_bounds(nt_total, world, r)[0] for r in range(world)]
        full = [torch.empty(n, 64, dtype=torch.uint8, device=device) for n in sizes]
        dist.all_gather(full, tl)
        tfull = torch.cat(full)
        exact = True
        for mode, cols in ((ab.MATCH_KNN2, 4), (ab.MATCH_COMPAT, 2)):
            a = ctx.match_sharded(q, tl, lo, mode).clone()
            b = ctx.match(q, tfull, mode)
            ctx.sync()
            exact = exact and bool(torch.equal(a[:, :cols], b[:, :cols]))
        flag = torch.tensor([1 if exact else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1, "train-sharded matching differs from the unsharded result"
        out["knn2_10kx1M_sharded_equals_unsharded"] = True
        out["knn2_10kx1M_how"] = "akz_match_sharded: akz_match(finalize=0) -> ncclAllGather(nq x 16 B per rank) on the library stream -> k_match_merge; no host sync"
        del full, tfull
        ctx.comm_destroy()
    else:
        ms = time_it(lambda: ctx.match(q, tl, ab.MATCH_KNN2, out=res), 3)
    out["knn2_10kx1M_ms"] = round(ms, 3)
    out["knn2_10kx1M_shards"] = world
    out["knn2_10kx1M_frac_of_plain_popc_bound"] = round(10000 * nt_total / (ms * 1e-3) / (peak_pairs * world), 3)
    out["knn2_10kx1M_frac_of_i8_tensor_peak"] = round(10000 * nt_total * 512 / (ms * 1e-3) / (peak_macs * world), 3)
    ctx.close()
    return out


def class_times(ctx, dev, res, F, chunk):
    """per-class device time: one extra pass with event pairs around every kernel group, chunk by chunk (so that with two lanes
    the kernels of this pass do not overlap each other: a kernel's time is its own)"""
    ctx.profile(True)
    for f0 in range(0, F, chunk):
        ctx.detect_and_compute(dev[f0:f0 + chunk], True, out=tuple(r[f0:f0 + chunk] for r in res))
    prof = ctx.profile_read()
    class_times.octaves = ctx.profile_octaves(8)          # the same pass split by octave (scale-space classes)
    ctx.profile(False)
    return prof


def noise_workload(ab, device, local, world, args):
    """SURVEY 8(d)'s primary input: uniform noise blurred with sigma 2 (~21 k keypoints per frame), max_pts 32768."""
    import torch
    F, mp = args.noise_frames, 32768
    frames8 = make_frames(F, "noise", seed0=0)
    dev = (torch.from_numpy(frames8).to(device).float() * (1.0 / 255.0)).contiguous()
    chunk = min(args.chunk, max(1, F // 2))               # two chunks at least: both lanes work
    ctx = ab.Context(W, H, max_batch=chunk, max_pts=mp, device=local, lanes=args.lanes)
    res = ctx.alloc_results(F, True)
    step = lambda: ctx.detect_and_compute(dev, True, out=res)
    for _ in range(3):
        step()
    ctx.sync()
    steps = max(4, args.steps)                             # 64 frames take 17 ms: a two-step region is at the mercy of one hiccup
    ms = timed(step, steps, ctx.torch_stream(), world, device)
    counts = res[0].cpu().numpy()
    prof = class_times(ctx, dev, res, F, chunk)
    tot = sum(v[0] for v in prof.values())
    out = {"value": round(F * world * steps / (ms * 1e-3), 2), "unit": "images/s", "frames_per_gpu": F, "chunk": chunk, "max_pts": mp,
           "keypoints_per_frame_mean": round(float(counts.mean()), 1), "clipped_frames": int((counts >= mp).sum()),
           "workload": "uniform u8 noise, Gaussian sigma 2, min-max stretched (tests/bindings.synth_noise_u8, seeds 0..7, rolled)",
           "classes_ms_per_step": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
           "keypoint_stage_share": round((prof.get("describe", (0, 0))[0] + prof.get("orient", (0, 0))[0]) / tot, 3) if tot else None}
    ctx.close()
    return out


def single_frame(device):
    """Batch-1 latency of the drop-in C++ surface (include/akaze.h through tests/cpp/dropin_shim.cu), as main.cpp:199-205 times it."""
    import torch
    import bindings as B
    if not B.have_dropin():
        return {"unavailable": "tests/cpp/build/libdropin_shim.so not built"}
    img = B.u8_to_unit(B.synth_shapes_u8(W, H, seed=1))
    pitch = (W + 127) // 128 * 128
    buf = np.zeros((H, pitch), dtype=np.float32)
    buf[:, :W] = img
    t = torch.from_numpy(buf).to(device)
    data = B.DropinData(10000)
    az = B.DropinAkazer(W, H, pitch)
    az.detect_and_compute(t, data)
    az.time(t, data, iters=5)
    ms = az.time(t, data, iters=50)
    az.time(t, data, iters=5, desc=False)             # the first two calls of an argument set run eagerly / capture the graph
    ms_nodesc = az.time(t, data, iters=50, desc=False)
    out = {"ms_per_frame": round(ms, 4), "detect_only_ms_per_frame": round(ms_nodesc, 4), "keypoints": int(data.num),
           "call": "akaze::Akazer::detectAndCompute(float*, AkazeData&, int3, true): device image in, AkazeData device + host records out, synchronous"}
    az.close()
    data.close()
    return out


def run_ours(args):
    import torch
    import akaze_b200 as ab
    rank, world, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    numa = bind_near_gpu(local, world)             # before the pinned allocations below
    F = args.frames
    frames8 = make_frames(F, args.content, seed0=100 * rank)
    dtype = np.float32 if args.dtype == "f32" else np.uint8
    host = torch.empty((F, H, W), dtype=torch.float32 if args.dtype == "f32" else torch.uint8).pin_memory()
    if args.dtype == "f32":
        host.numpy()[:] = frames8.astype(np.float32) * np.float32(1.0 / 255.0)       # main.cpp:149
    else:
        host.numpy()[:] = frames8
    ctx = ab.Context(W, H, max_batch=args.chunk, max_pts=args.max_pts, device=local, lanes=args.lanes)
    stream = ctx.torch_stream()
    dev = host.to(device)
    res = ctx.alloc_results(F, True)
    mp = args.max_pts
    h_counts = torch.zeros(F, dtype=torch.int32).pin_memory()
    h_kpts = torch.zeros((F, mp, 8), dtype=torch.int32).pin_memory()
    h_desc = torch.zeros((F, mp, 64), dtype=torch.uint8).pin_memory()
    hout = (h_counts.numpy(), h_kpts.numpy().view(ab.KEYPOINT_DTYPE).reshape(F, mp), h_desc.numpy())

    def step_dev():
        ctx.detect_and_compute(dev, True, out=res)

    def step_host():
        ctx.detect_and_compute_host(host, True, out=hout)

    for _ in range(args.warmup):
        step_dev()
    ctx.sync()
    l0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps, stream, world, device)
    launches = ctx.launches - l0
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    ms_e2e = timed(step_host, args.steps, stream, world, device)
    counts = res[0].cpu().numpy()
    assert np.array_equal(counts, h_counts.numpy()), "host and device paths disagree"
    # the same end-to-end call with raw 8-bit frames (the reference's fastDetectAndCompute ingest; a quarter of the H2D bytes)
    ms_e2e_u8 = None
    if args.dtype == "f32":
        host8 = torch.from_numpy(frames8).pin_memory()
        step8 = lambda: ctx.detect_and_compute_host(host8, True, out=hout)
        step8()
        ms_e2e_u8 = timed(step8, args.steps, stream, world, device)
        assert np.array_equal(counts, h_counts.numpy()), "u8 and f32 ingest disagree"
        del host8
    clocks = sampler.stop() if rank == 0 else None
    nkp_mean = float(counts.mean())

    prof = class_times(ctx, dev, res, F, args.chunk)
    tot_ms = sum(v[0] for v in prof.values())
    top = max(prof, key=lambda k: prof[k][0])
    top_ms, top_launches = prof[top]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = algorithmic_bytes(top, nkp_mean) * F
    achieved = alg / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:                                     # DRAM bytes of the captured launch, scaled by the algorithmic bytes (profiles/traffic.json)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if top in tr:
            traffic = tr[top]["ratio"] * alg / max(1, top_launches)
            traffic_src = f"ncu {tr[top]['kernel']}: {tr[top]['dram_bytes']} B DRAM for {tr[top]['algorithmic_bytes']} B algorithmic ({tr['source']})"
    except Exception:
        pass
    step_alg = (4 * W * H + 20 * sum(level_pixels(W, H))) * F          # 228.6 MB / frame
    roof = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": None if traffic is None else round(traffic), "traffic_source": traffic_src,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "kernel_ms_per_step": round(top_ms, 3), "kernel_share_of_step": round(top_ms / tot_ms, 3) if tot_ms else None,
            "kernel_launches_per_step": int(top_launches), "algorithmic_bytes_per_launch": alg / max(1, top_launches),
            "classes_ms_per_step": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "whole_step": {"algorithmic_gbs": round(step_alg / (ms / args.steps * 1e-3) / 1e9, 1),
                           "frac_of_hbm": round(step_alg / (ms / args.steps * 1e-3) / 1e9 / peak, 4)}}
    # the scale-space classes by octave: octave 0 holds 75 % of the pixels and is where the streaming kernels run full waves;
    # octaves 1-3 are small launches that the second lane and the octave streams overlap in the real step
    px = level_pixels(W, H)
    by_oct = {}
    for cls in ("prep", "fed"):
        row = getattr(class_times, "octaves", {}).get(cls)
        if not row:
            continue
        out_rows = []
        for o in range(len(px) // 4):
            lv = px[4 * o:4 * o + 4]
            # prep: 20 B/px per level (read Lt_prev; write g, Lx, Ly, det), level (0,0) has no g (16), octave transitions also write Lt (24)
            b = (20 * sum(lv) - (4 * lv[0] if o == 0 else -4 * lv[0])) if cls == "prep" else 12 * (sum(lv) - (lv[0] if o == 0 else 0))
            if row[o] > 0:
                gbs = b * F / (row[o] * 1e-3) / 1e9
                out_rows.append({"octave": o, "ms_per_step": round(row[o], 3), "algorithmic_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
        by_oct[cls] = out_rows
    roof["scale_space_by_octave"] = by_oct
    line = {
        "metric": "1080p detect+describe images/sec", "value": round(F * world * args.steps / (ms * 1e-3), 2), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"configs[2]: synthetic 1920x1080 grayscale, {F} frames per GPU per step ({args.content}), "
                               "4 octaves x 4 sublevels, reference defaults (main.cpp:156-166), detect+describe",
                   "frames_per_gpu": F, "chunk": args.chunk, "lanes": args.lanes, "max_pts": mp, "keypoints_per_frame_mean": round(nkp_mean, 1),
                   "parallelism": f"frame-sharded x{world}, no collective",
                   "l2_policy": f"inputs larger than L2: {F} frames x {W * H * (4 if args.dtype == 'f32' else 1) / 1e6:.1f} MB in, "
                                f"{args.chunk} x 176 MB of planes per chunk (L2 = 126 MB)",
                   "roofline_pass": "one extra pass of the same step with CUDA event pairs around every kernel group"},
        "e2e": {"value": round(F * world * args.steps / (ms_e2e * 1e-3), 2), "unit": "images/s",
                "h2d_bytes_per_step": int(host.numel() * host.element_size()),
                "d2h_bytes_per_step": int(F * 4 + counts.sum() * (32 + 64)), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "e2e_u8": None if ms_e2e_u8 is None else {"value": round(F * world * args.steps / (ms_e2e_u8 * 1e-3), 2), "unit": "images/s",
                                                  "h2d_bytes_per_step": int(F * W * H), "note": "same call with raw u8 host frames (AKZ_U8)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
    }
    ctx.close()
    # what the same kernels cost on a FULL machine: the profile pass again with one lane and chunks of 128 frames (the real step
    # gets the same effect from its second lane and the octave streams: its time does not change with the chunk size)
    if rank == 0 and world == 1 and F >= 128 and not args.no_full:
        try:
            big = ab.Context(W, H, max_batch=128, max_pts=mp, device=local, lanes=1)
            for _ in range(2):
                big.detect_and_compute(dev[:128], True, out=tuple(r[:128] for r in res))
            big.sync()
            Fb = F - F % 128
            pf = class_times(big, dev[:Fb], tuple(r[:Fb] for r in res), Fb, 128)
            big.close()
            sc = F / Fb
            alg_p = algorithmic_bytes("prep", nkp_mean) * F
            alg_f = algorithmic_bytes("fed", nkp_mean) * F
            roof["full_machine"] = {"chunk": 128, "lanes": 1,
                                    "classes_ms_per_step": {k: round(v[0] * sc, 3) for k, v in sorted(pf.items(), key=lambda kv: -kv[1][0])},
                                    "prep_frac": round(alg_p / (pf["prep"][0] * sc * 1e-3) / 1e9 / peak, 4),
                                    "fed_frac": round(alg_f / (pf["fed"][0] * sc * 1e-3) / 1e9 / peak, 4)}
        except Exception as e:                                   # never lose the line over the extra pass
            roof["full_machine"] = {"unavailable": str(e)[:200]}
    del dev
    line["config"]["binding"] = numa
    line["e2e"]["h2d_gbs_per_rank"] = round(host.numel() * host.element_size() / (ms_e2e / args.steps * 1e-3) / 1e9, 2)
    del host
    if not args.no_noise:
        line["noise"] = noise_workload(ab, device, local, world, args)
    if rank == 0 and world == 1:
        line["single_frame"] = single_frame(device)
    if not args.no_match:
        m = match_metric(ab, device, rank, world, (clocks or {}).get("sm_mhz") if clocks else None)
        line["match"] = m
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(frames8, args.cpu_frames)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(frames8, nsample):
    """OpenCV cv::AKAZE (what main.cpp:344-399 times as the CPU arm) and the C oracle port on the first frames."""
    cores = os.cpu_count() or 1
    out = {"unit": "images/s", "cores": cores}
    sample = frames8[:nsample]
    try:
        import cv2
        cv2.setNumThreads(cores)
        ak = cv2.AKAZE_create()
        ak.detectAndCompute(sample[0], None)
        t0 = time.perf_counter()
        nk = 0
        for f in sample:
            kp, _ = ak.detectAndCompute(f, None)
            nk += len(kp)
        dt = time.perf_counter() - t0
        out.update({"value": round(len(sample) / dt, 3), "kind": "reference",
                    "sample": f"cv2 {cv2.__version__} AKAZE_create() defaults (main.cpp:373) on the first {len(sample)} frames of the batch, "
                              f"{cv2.getNumThreads()} threads, {nk / len(sample):.0f} keypoints/frame"})
    except Exception as e:                                     # cv2 missing on the box: the port below is the baseline
        out["opencv_error"] = repr(e)[:120]
    try:
        import bindings as B
        img = B.u8_to_unit(sample[0])
        t0 = time.perf_counter()
        k = B.oracle_detect_and_compute(img, True, threads=cores)
        dt = time.perf_counter() - t0
        port = {"value": round(1.0 / dt, 3), "kind": "port", "cores": cores,
                "sample": f"oracle/akaze_oracle.c (OpenMP, {cores} threads) on the first frame, {len(k)} keypoints"}
        if "value" in out:
            out["port"] = port
        else:
            out.update(port)
    except Exception as e:
        out["port_error"] = repr(e)[:120]
    # the CPU matcher beside metric 2, and configs[0] (left.pgm + right.pgm through OpenCV: detect, describe, match)
    try:
        import cv2
        import bindings as B
        q, t = B.random_descriptors(10000, 0)[:, :61].copy(), B.random_descriptors(10000, 1)[:, :61].copy()
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        t0 = time.perf_counter()
        mm = bf.knnMatch(q, t, k=2)
        out["bfmatcher_knn2_10kx10k_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        out["bfmatcher_note"] = f"cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2), 61-byte rows, {cv2.getNumThreads()} threads, {len(mm)} queries"
        l, r = os.path.join(B.REF_DATA, "left.pgm"), os.path.join(B.REF_DATA, "right.pgm")
        if os.path.exists(l) and os.path.exists(r):
            a, b = B.read_pgm(l), B.read_pgm(r)
            ak = cv2.AKAZE_create()
            t0 = time.perf_counter()
            k1, d1 = ak.detectAndCompute(a, None)
            k2, d2 = ak.detectAndCompute(b, None)
            t1 = time.perf_counter()
            m = bf.match(d1, d2)
            t2 = time.perf_counter()
            out["configs0_left_right_pgm"] = {"detect_describe_ms_per_image": round((t1 - t0) * 500, 1), "match_ms": round((t2 - t1) * 1e3, 1),
                                              "keypoints": [len(k1), len(k2)], "matches": len(m),
                                              "how": "cv2.AKAZE_create() defaults + BFMatcher(NORM_HAMMING).match, 1280x960 u8 (main.cpp:344-399)"}
    except Exception as e:
        out["cpu_match_error"] = repr(e)[:120]
    return out


def run_reference(args):
    """The unmodified reference CUDA library through Akazer::detectAndCompute, one frame per call (main.cpp:199-205)."""
    import torch
    rank, world, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    import bindings as B
    if not B.have_ref():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_akaze.so not built (needs /root/reference at build time)"}))
        return
    F = args.frames
    frames8 = make_frames(F, args.content, seed0=100 * rank)
    pitch = (W + 127) // 128 * 128                                  # main.cpp:174: iAlignUp(w, 128)
    host = torch.zeros((F, H, pitch), dtype=torch.float32).pin_memory()
    host.numpy()[:, :, :W] = frames8.astype(np.float32) * np.float32(1.0 / 255.0)
    dev = host.to(device)
    mp = args.max_pts
    ref = B.RefAkazer(W, H, pitch)
    L = ref.L
    import ctypes as C
    d_pts = torch.zeros(mp * 104, dtype=torch.uint8, device=device)
    h_pts = torch.zeros(mp * 104, dtype=torch.uint8).pin_memory()
    stage = torch.empty((H, pitch), dtype=torch.float32, device=device)
    counts = np.zeros(F, dtype=np.int64)

    def step_dev():
        for f in range(F):
            counts[f] = L.ref_akazer_detectAndCompute(ref.hnd, C.c_void_p(dev[f].data_ptr()), W, H, pitch, 1,
                                                      C.c_void_p(d_pts.data_ptr()), None, mp)

    def step_host():
        for f in range(F):
            stage.copy_(host[f], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            counts[f] = L.ref_akazer_detectAndCompute(ref.hnd, C.c_void_p(stage.data_ptr()), W, H, pitch, 1,
                                                      C.c_void_p(d_pts.data_ptr()), C.c_void_p(h_pts.data_ptr()), mp)

    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps, stream, world, device)
    ms_e2e = timed(step_host, args.steps, stream, world, device)
    clocks = sampler.stop() if rank == 0 else None
    v = F * world * args.steps / (ms * 1e-3)
    # metric 2 with the reference's own matcher (gHammingMatch through hMatch): 10k x 10k AkazePoint records.  The train
    # count must be a multiple of 16: the kernel deadlocks on sm_100a otherwise (akazed.cu:2176-2187, App. B-15).
    match = None
    if not args.no_match:
        q = B.random_descriptors(10000, 0); t = B.random_descriptors(10000, 1)
        pq = np.zeros(10000, dtype=B.REF_POINT); pt = np.zeros(10000, dtype=B.REF_POINT)
        pq["features"], pt["features"] = q[:, :61], t[:, :61]
        dq = torch.from_numpy(pq.view(np.uint8).reshape(-1)).to(device); dt = torch.from_numpy(pt.view(np.uint8).reshape(-1)).to(device)
        torch.cuda.synchronize()
        for _ in range(2):
            L.ref_hMatch(C.c_void_p(dq.data_ptr()), 10000, C.c_void_p(dt.data_ptr()), 10000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.ref_hMatch(C.c_void_p(dq.data_ptr()), 10000, C.c_void_p(dt.data_ptr()), 10000)
        e1.record(); e1.synchronize()
        match = {"compat_10kx10k_ms": round(e0.elapsed_time(e1) / 5, 4), "how": "akaze::hMatch (gHammingMatch, 16 threads per query), 1-NN with the uniqueness gate"}
    # the keypoint-heavy input and the single-frame latency, same definitions as in the main arm
    noise = None
    if not args.no_noise:
        Fn = min(16, args.noise_frames)
        n8 = make_frames(Fn, "noise", seed0=0)
        hn = torch.zeros((Fn, H, pitch), dtype=torch.float32)
        hn.numpy()[:, :, :W] = n8.astype(np.float32) * np.float32(1.0 / 255.0)
        dn = hn.to(device)
        npts = torch.zeros(32768 * 104, dtype=torch.uint8, device=device)
        ncnt = np.zeros(Fn, dtype=np.int64)

        def nstep():
            for f in range(Fn):
                ncnt[f] = L.ref_akazer_detectAndCompute(ref.hnd, C.c_void_p(dn[f].data_ptr()), W, H, pitch, 1, C.c_void_p(npts.data_ptr()), None, 32768)
        nstep()
        nms = timed(nstep, 2, stream, world, device)
        noise = {"value": round(Fn * world * 2 / (nms * 1e-3), 2), "unit": "images/s", "frames_per_gpu": Fn, "max_pts": 32768,
                 "keypoints_per_frame_mean": round(float(ncnt.mean()), 1)}
        del dn, npts
    sf = None
    if rank == 0 and world == 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            L.ref_akazer_detectAndCompute(ref.hnd, C.c_void_p(dev[0].data_ptr()), W, H, pitch, 1, C.c_void_p(d_pts.data_ptr()), C.c_void_p(h_pts.data_ptr()), mp)
        torch.cuda.synchronize()
        sf = {"ms_per_frame": round((time.perf_counter() - t0) * 50.0, 4), "call": "Akazer::detectAndCompute of the reference, device image in, host records out"}
    line = {
        "impl": "reference", "match": match, "noise": noise, "single_frame": sf, "metric": "1080p detect+describe images/sec", "value": round(v, 2), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2]: synthetic 1920x1080 grayscale, {F} frames per GPU per step ({args.content}), "
                               "4 octaves x 4 sublevels, reference defaults (main.cpp:156-166), detect+describe",
                   "frames_per_gpu": F, "max_pts": mp, "keypoints_per_frame_mean": round(float(counts.mean()), 1),
                   "how": "unmodified akazed.cu/akaze.cpp/fed.cpp compiled for sm_100a (oracle/Makefile), Akazer::detectAndCompute per frame "
                          "on the reused-size path, legacy default stream; the reference is a CUDA library and has no CPU implementation"},
        "e2e": {"value": round(F * world * args.steps / (ms_e2e * 1e-3), 2), "unit": "images/s",
                "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": int(counts.sum() * 85),
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "cpu_baseline": {"value": None, "unit": "images/s", "cores": 0, "kind": "reference",
                         "sample": "not a CPU run: the reference implementation of this path is CUDA; its CPU comparison (cv::AKAZE) is in the main arm's cpu_baseline"},
        "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    ref.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step (configs[2]: 256)")
    ap.add_argument("--chunk", type=int, default=64, help="frames processed together (akz_options.max_batch); the host path caps the chunks "
                    "of float frames at 32 (link-bound: fill / drain), see akz_detect_and_compute_host")
    ap.add_argument("--lanes", type=int, default=2, help="chunks in flight (akz_options.lanes): 2 = two streams with their own pyramids")
    ap.add_argument("--max-pts", type=int, default=10000, help="per-frame keypoint capacity (main.cpp:157)")
    ap.add_argument("--content", default="shapes", choices=["shapes", "noise"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "u8"])
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-match", action="store_true", help="skip the brute-force matching metric (BASELINE metric 2)")
    ap.add_argument("--no-noise", action="store_true", help="skip the keypoint-heavy noise workload")
    ap.add_argument("--no-full", action="store_true", help="skip the extra profile pass with chunks of 128 frames (roofline.full_machine)")
    ap.add_argument("--noise-frames", type=int, default=64, help="frames per GPU of the noise workload")
    ap.add_argument("--config", default="frames", choices=["frames", "stream"], help="frames = configs[2] (default), stream = configs[4]")
    ap.add_argument("--stream-frames", type=int, default=32, help="configs[4]: frames of the stream per GPU (plus the overlap frame)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)
    if args.config == "stream":
        import bench_stream
        bench_stream.run(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
