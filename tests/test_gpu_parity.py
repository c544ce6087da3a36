"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test drives the product through the C ABI (akaze_b200.Context -> libakaze_b200.so) and checks it against
  * the reference itself, compiled for sm_100a (oracle/_ref/libref_akaze.so, bindings.ref()), and
  * the CPU restatement (oracle/liboracle_akaze.so),
on identical inputs.  Bars: planes, descriptors and match indices bit-exact; refined positions
<= 1e-4 px (north-star bar 0.01 px); orientation <= 1e-4 rad.  Sizes are the ones at which the reference itself is
well defined (App. B-7: 1280x960 and heights that are multiples of 16)."""
import ctypes as C
import os

import numpy as np
import pytest

import bindings as B

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def ab():
    import akaze_b200
    return akaze_b200


def dev(a, pitch=None):
    a = np.ascontiguousarray(a)
    h, w = a.shape
    pitch = pitch or w
    buf = np.zeros((1, h, pitch), dtype=a.dtype)
    buf[0, :, :w] = a
    t = torch.from_numpy(buf).cuda()
    torch.cuda.synchronize()
    return t


def host(t, w, frame=0):
    torch.cuda.synchronize()
    return t[frame, :, :w].cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, name):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    bad = np.argwhere(a != b)
    if len(bad):
        y, x = bad[0]
        fa, fb = a.view(np.float32), b.view(np.float32)
        raise AssertionError(f"{name}: {len(bad)} of {a.size} values differ; first at (y={y}, x={x}): "
                             f"{fa[y, x]!r} vs {fb[y, x]!r}; max abs diff {np.abs(fa - fb).max():.3e}")


def left_image():
    p = os.path.join(B.REF_DATA, "left.pgm")
    if os.path.exists(p):
        return B.u8_to_unit(B.read_pgm(p))
    return B.u8_to_unit(B.synth_shapes_u8(1280, 960, seed=7))


def right_image():
    p = os.path.join(B.REF_DATA, "right.pgm")
    if os.path.exists(p):
        return B.u8_to_unit(B.read_pgm(p))
    return B.u8_to_unit(B.synth_shapes_u8(1280, 960, seed=8))


needs_ref = pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref/libref_akaze.so not built")


@pytest.fixture(scope="module")
def stage_ctx():
    c = ab().Context(0, 0, fused=0, max_batch=4)
    yield c
    c.close()


@pytest.fixture(scope="module")
def stage_ctx_fused():
    c = ab().Context(0, 0, fused=1, max_batch=4)
    yield c
    c.close()


# ---------------------------------------------------------------------------------------------------------------
# stage seams vs the reference's own stage functions (akazed.h) and the CPU oracle
# ---------------------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("var,ksz", [(1.0, 5), (2.56, 9), (1.5, 7), (3.0, 11)])
def test_lowpass_vs_reference(stage_ctx, var, ksz):
    img = left_image()
    h, w = img.shape
    src = dev(img)
    mine = torch.zeros_like(src)
    refo = torch.zeros_like(src)
    stage_ctx.lowpass(src, mine, w, var, ksz)
    stage_ctx.sync()
    B.ref().ref_hLowPass(C.c_void_p(src.data_ptr()), C.c_void_p(refo.data_ptr()), w, h, w, var, ksz)
    assert_bits_equal(host(mine, w), host(refo, w), f"lowpass var={var} vs reference")
    orc = np.zeros_like(img)
    B.oracle().orc_lowpass(B._p(np.ascontiguousarray(img)), B._p(orc), w, h, w, var, ksz)
    assert_bits_equal(host(mine, w), orc, f"lowpass var={var} vs CPU oracle")


@needs_ref
def test_down_with_smooth_vs_reference(stage_ctx):
    img = left_image()
    h, w = img.shape
    dw, dh = w >> 1, h >> 1
    src = dev(img)
    d1, s1 = torch.zeros(1, dh, dw, device="cuda"), torch.zeros(1, dh, dw, device="cuda")
    d2, s2 = torch.zeros_like(d1), torch.zeros_like(s1)
    stage_ctx.down_with_smooth(src, w, d1, s1, dw)
    stage_ctx.sync()
    B.ref().ref_hDownWithSmooth(C.c_void_p(src.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_void_p(s2.data_ptr()), w, h, w, dw, dh, dw)
    assert_bits_equal(host(d1, dw), host(d2, dw), "downsample vs reference")
    assert_bits_equal(host(s1, dw), host(s2, dw), "coarse-lattice blur vs reference")
    od, os_ = np.zeros((dh, dw), np.float32), np.zeros((dh, dw), np.float32)
    B.oracle().orc_down_with_smooth(B._p(np.ascontiguousarray(img)), B._p(od), B._p(os_), w, h, w, dw, dh, dw)
    assert_bits_equal(host(s1, dw), os_, "coarse-lattice blur vs CPU oracle")


def test_down_with_smooth_odd_sizes_vs_oracle(stage_ctx):
    for (w, h) in [(135, 101), (240, 135), (97, 64)]:
        img = B.u8_to_unit(B.synth_noise_u8(w, h, seed=3))
        dw, dh = w >> 1, h >> 1
        dp = (dw + 31) // 32 * 32
        src = dev(img)
        d1, s1 = torch.zeros(1, dh, dp, device="cuda"), torch.zeros(1, dh, dp, device="cuda")
        stage_ctx.down_with_smooth(src, w, d1, s1, dw)
        stage_ctx.sync()
        od, os_ = np.zeros((dh, dw), np.float32), np.zeros((dh, dw), np.float32)
        B.oracle().orc_down_with_smooth(B._p(np.ascontiguousarray(img)), B._p(od), B._p(os_), w, h, w, dw, dh, dw)
        assert_bits_equal(host(d1, dw), od, f"downsample {w}x{h}")
        assert_bits_equal(host(s1, dw), os_, f"coarse blur {w}x{h}")


@needs_ref
@pytest.mark.parametrize("dtype_k", [(1, 0.0123), (1, 0.0021), (0, 0.02), (3, 0.02), (2, 0.02)])
def test_flow_vs_reference(stage_ctx, dtype_k):
    typ, k = dtype_k
    img = left_image()
    h, w = img.shape
    sm = np.zeros_like(img)
    B.oracle().orc_lowpass(B._p(np.ascontiguousarray(img)), B._p(sm), w, h, w, 1.0, 5)
    src = dev(sm)
    f1, f2 = torch.zeros_like(src), torch.zeros_like(src)
    kt = torch.tensor([k], dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    stage_ctx.flow(src, f1, w, kt, type=typ)
    stage_ctx.sync()
    B.ref().ref_hFlow(C.c_void_p(src.data_ptr()), C.c_void_p(f2.data_ptr()), typ, k, w, h, w)
    assert_bits_equal(host(f1, w), host(f2, w), f"conductance type {typ} vs reference")
    if typ == 1:
        of = np.zeros_like(sm)
        B.oracle().orc_flow(B._p(sm), B._p(of), typ, k, w, h, w)
        assert_bits_equal(host(f1, w), of, "conductance vs CPU oracle")


@needs_ref
def test_nld_step_vs_reference(stage_ctx):
    img = left_image()
    h, w = img.shape
    sm = np.zeros_like(img)
    fl = np.zeros_like(img)
    B.oracle().orc_lowpass(B._p(np.ascontiguousarray(img)), B._p(sm), w, h, w, 1.0, 5)
    B.oracle().orc_flow(B._p(sm), B._p(fl), 1, 0.01, w, h, w)
    L, g = dev(img), dev(fl)
    o1, o2 = torch.zeros_like(L), torch.zeros_like(L)
    for tau in (0.0697, 0.35204, 41.3):
        stage_ctx.nld_step(L, g, o1, w, tau)
        stage_ctx.sync()
        B.ref().ref_hNldStep(C.c_void_p(L.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(o2.data_ptr()), tau, w, h, w)
        assert_bits_equal(host(o1, w), host(o2, w), f"NLD step tau={tau} vs reference")
        oo = np.zeros_like(img)
        B.oracle().orc_nld_step(B._p(np.ascontiguousarray(img)), B._p(fl), B._p(oo), tau, w, h, w)
        assert_bits_equal(host(o1, w), oo, f"NLD step tau={tau} vs CPU oracle")


@needs_ref
@pytest.mark.parametrize("step", [2, 3, 4])
def test_hessian_vs_reference(stage_ctx, stage_ctx_fused, step):
    img = left_image()
    h, w = img.shape
    sm = np.zeros_like(img)
    B.oracle().orc_lowpass(B._p(np.ascontiguousarray(img)), B._p(sm), w, h, w, 1.0, 5)
    src = dev(sm)
    outs = {}
    for name, ctx in (("stage", stage_ctx), ("fused", stage_ctx_fused)):
        lx, ly, det = torch.zeros_like(src), torch.zeros_like(src), torch.zeros_like(src)
        ctx.hessian(src, lx, ly, det, w, step)
        ctx.sync()
        outs[name] = (host(lx, w), host(ly, w), host(det, w))
    rsrc = src.clone()
    rlx, rly = torch.zeros_like(src), torch.zeros_like(src)
    torch.cuda.synchronize()
    B.ref().ref_hHessianDeterminant(C.c_void_p(rsrc.data_ptr()), C.c_void_p(rlx.data_ptr()), C.c_void_p(rly.data_ptr()), step, w, h, w)
    refv = (host(rlx, w), host(rly, w), host(rsrc, w))
    olx, oly, odet = np.zeros_like(sm), np.zeros_like(sm), np.zeros_like(sm)
    B.oracle().orc_hessian(B._p(sm), B._p(olx), B._p(oly), B._p(odet), step, w, h, w)
    for i, nm in enumerate(("Lx", "Ly", "det")):
        assert_bits_equal(outs["stage"][i], refv[i], f"{nm} step={step} stage kernels vs reference")
        assert_bits_equal(outs["fused"][i], refv[i], f"{nm} step={step} fused kernel vs reference")
        assert_bits_equal(outs["stage"][i], (olx, oly, odet)[i], f"{nm} step={step} vs CPU oracle")


@needs_ref
def test_contrast_factor(stage_ctx):
    """True-maximum contrast factor: equals the CPU oracle exactly; the reference's own value comes from a racy
    reduction (App. B-1) and is reported, not required."""
    img = left_image()
    h, w = img.shape
    sm = np.zeros_like(img)
    B.oracle().orc_lowpass(B._p(np.ascontiguousarray(img)), B._p(sm), w, h, w, 1.0, 5)
    src = dev(sm)
    k = float(stage_ctx.scharr_contrast(src, w, 0.7).cpu()[0])
    mag = np.zeros_like(sm)
    B.oracle().orc_scharr_mag(B._p(sm), B._p(mag), w, h, w)
    hmax = C.c_float()
    ko = B.oracle().orc_contrast_from_mag(B._p(mag), w, h, w, 0.7, C.byref(hmax))
    assert np.float32(k).view(np.uint32) == np.float32(ko).view(np.uint32), (k, ko)
    grad = torch.zeros_like(src)
    kr = [B.ref().ref_hScharrContrast(C.c_void_p(src.data_ptr()), C.c_void_p(grad.data_ptr()), 0.7, w, h, w) for _ in range(3)]
    print(f"\n[contrast] ours={k:.7f} oracle={ko:.7f} hmax={hmax.value:.6f} reference runs={kr}")
    assert abs(k - kr[0]) <= 0.25 * k        # same ballpark; the reference's maximum is a strided subsample


# ---------------------------------------------------------------------------------------------------------------
# fused kernels vs per-stage kernels
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 3, 7, 8, 9, 14, 29])
@pytest.mark.parametrize("size", [(333, 217, 32), (1280, 96, 32), (64, 300, 32), (121, 70, 4), (480, 270, 32), (250, 33, 1), (333, 100, 1)])
def test_fed_cycle_fused_equals_single_steps(stage_ctx, stage_ctx_fused, n, size):
    """The FED cycle kernels against n single-step launches and the CPU oracle.  Row pitches that are multiples of 4 floats run
    the streaming warp kernel k_fed4 (widths that are not multiples of 4 its general variant; 121 = one full strip of 120
    columns plus a strip of one column); unaligned pitches (last entries) run the tile kernel k_fed3."""
    w, h, align = size
    p = (w + align - 1) // align * align
    rng = np.random.default_rng(n)
    L = rng.random((2, h, p), dtype=np.float32)
    g = rng.random((2, h, p), dtype=np.float32)
    tau = (rng.random(n, dtype=np.float32) * 0.2 + 0.01).astype(np.float32)
    Lt, gt = torch.from_numpy(L).cuda(), torch.from_numpy(g).cuda()
    outs = []
    for ctx in (stage_ctx, stage_ctx_fused):
        dst, tmp = torch.zeros_like(Lt), torch.zeros_like(Lt)
        torch.cuda.synchronize()
        ctx.fed_cycle(Lt, gt, dst, tmp, w, tau)
        ctx.sync()
        outs.append(dst.cpu().numpy()[:, :, :w])
    for f in range(2):
        assert_bits_equal(outs[0][f], outs[1][f], f"FED n={n} {w}x{h} frame {f}: fused vs single steps")
    # and against the CPU oracle for frame 0
    cur = np.ascontiguousarray(L[0])
    for k in range(n):
        nxt = np.zeros_like(cur)
        B.oracle().orc_nld_step(B._p(cur), B._p(np.ascontiguousarray(g[0])), B._p(nxt), float(tau[k]), w, h, p)
        cur = nxt
    assert_bits_equal(outs[1][0], cur[:, :w], f"FED n={n} {w}x{h}: fused vs CPU oracle")


def _planes(ctx, frame=0):
    out = []
    for l in range(ctx.num_levels):
        out.append([ctx.plane(l, which, frame) for which in range(4)])
    return out


@pytest.mark.parametrize("size", [(1280, 960), (333, 250), (640, 480)])
def test_scale_space_fused_equals_stages_equals_oracle(size):
    w, h = size
    img8 = B.synth_shapes_u8(w, h, seed=11) if size != (1280, 960) else None
    img = left_image() if img8 is None else B.u8_to_unit(img8)
    res = {}
    for fused in (0, 1):
        ctx = ab().Context(w, h, fused=fused, max_batch=2, max_pts=20000)
        t = torch.from_numpy(np.stack([img, img[::-1].copy()])).cuda()
        torch.cuda.synchronize()
        ctx.build_scale_space(t)
        ctx.sync()
        res[fused] = (_planes(ctx, 0), _planes(ctx, 1), ctx.kcontrast(2))
        ctx.close()
    names = ["Lt", "det", "Lx", "Ly"]
    for f in range(2):
        for l, (a, b) in enumerate(zip(res[0][f], res[1][f])):
            for which in range(4):
                assert_bits_equal(a[which], b[which], f"{w}x{h} frame {f} level {l} {names[which]}: stages vs fused")
    assert np.array_equal(res[0][2].view(np.uint32), res[1][2].view(np.uint32))
    # CPU oracle on frame 0
    P = B.OraclePyramid(w, h)
    P.build(img)
    assert np.float32(P.kcontrast).view(np.uint32) == res[1][2][:1].view(np.uint32)[0], (P.kcontrast, res[1][2])
    for l in range(P.levels):
        for which in range(4):
            assert_bits_equal(res[1][0][l][which], P.plane(l, which), f"{w}x{h} level {l} {names[which]}: fused vs CPU oracle")
    P.close()


# ---------------------------------------------------------------------------------------------------------------
# whole pipeline vs the reference
# ---------------------------------------------------------------------------------------------------------------
def _kp_array(counts, kpts, frame=0):
    torch.cuda.synchronize()
    n = int(counts[frame].cpu())
    return ab().keypoints_from_words(kpts[frame, :n].cpu().numpy())


@pytest.fixture(scope="module")
def ref_left():
    """One run of the reference pipeline on left.pgm, planes kept."""
    if not B.have_ref():
        pytest.skip("reference not built")
    img = left_image()
    h, w = img.shape
    r = B.RefAkazer(w, h, w)
    pts, planes, k = r.detect_keep(dev(img)[0], max_pts=30000)
    r.close()
    return dict(img=img, pts=pts, planes=planes, k=k)


def test_scale_space_vs_reference(ref_left):
    img = ref_left["img"]
    h, w = img.shape
    names = ["Lt", "det", "Lx", "Ly"]
    for fused in (1, 0):
        ctx = ab().Context(w, h, fused=fused, max_batch=1, max_pts=30000, kcontrast_override=ref_left["k"])
        ctx.build_scale_space(dev(img))
        ctx.sync()
        mine = _planes(ctx)
        ctx.close()
        assert len(mine) == len(ref_left["planes"])
        for l in range(len(mine)):
            for which in range(4):
                assert_bits_equal(mine[l][which], ref_left["planes"][l][which], f"level {l} {names[which]} (fused={fused}) vs reference")


def _match_sets(mine, refp):
    km = {(int(k["layer"]), int(k["iy"]), int(k["ix"])): i for i, k in enumerate(mine)}
    return km


def test_keypoints_vs_reference(ref_left):
    img = ref_left["img"]
    h, w = img.shape
    ctx = ab().Context(w, h, max_batch=1, max_pts=30000, kcontrast_override=ref_left["k"])
    counts, kpts, desc = ctx.detect_and_compute(dev(img))
    ctx.sync()
    mine = _kp_array(counts, kpts)
    torch.cuda.synchronize()
    dm = desc[0].cpu().numpy()
    ctx.close()
    from scipy.spatial import cKDTree
    tree = cKDTree(np.stack([mine["x"], mine["y"]], 1))

    def compare(refp):
        """Returns the list of statistical bars this reference run misses (the reference's sublevel merge is a live data race,
        App. B-2: its outcome differs from run to run); the descriptor comparison is exact and asserted outright."""
        # reference integer positions are not stored after refinement; match on layer + nearest refined position
        d, j = tree.query(np.stack([refp["x"], refp["y"]], 1))
        same = (d <= 1e-4) & (mine["layer"][j] == refp["octave"])
        rep = same.mean()
        print(f"\n[keypoints] ours={len(mine)} reference={len(refp)} identical-position fraction={rep:.5f} "
              f"max position error among matched={d[same].max() if same.any() else -1:.2e}")
        missed = []
        # The whole-pipeline reference run merges sublevels with a data race that is live on a B200 at octaves >= 1 (all
        # z-slices of gCalcExtremaMap are co-resident): a racing pixel can keep the SMALLER response, which changes what the
        # radius NMS around it suppresses.  The exact, race-free comparison is test_detector_vs_serialized_reference (set
        # equality); here the racy run must still agree on the bulk.
        if abs(len(mine) - len(refp)) > 0.04 * len(refp):
            missed.append(f"count {len(mine)} vs {len(refp)}")
        if rep < 0.93:
            missed.append(f"identical-position fraction {rep:.4f}")
        sz = np.mean(mine["size"][j][same] == refp["size"][same])               # torn (layer, size) pairs of the racy merge
        if sz < 0.995:
            missed.append(f"size agreement {sz:.4f}")
        # orientation: <= 1e-4 rad modulo 2 pi (App. B-4: the reference sums its histogram with float atomics)
        da = np.abs(mine["angle"][j][same] - refp["angle"][same])
        da = np.minimum(da, 2 * np.pi - da)
        frac = (da <= 1e-4).mean()
        print(f"[orientation] max diff {da.max():.3e}, fraction within 1e-4 rad: {frac:.5f}")
        if frac < 0.995:
            missed.append(f"orientation agreement {frac:.4f}")
        # descriptors of keypoints whose position AND angle bits agree must be identical
        exact = same & (mine["angle"][j].view(np.uint32) == refp["angle"].view(np.uint32)) & \
            (mine["x"][j].view(np.uint32) == refp["x"].view(np.uint32)) & (mine["y"][j].view(np.uint32) == refp["y"].view(np.uint32))
        ours = dm[j[exact]][:, :61]
        theirs = refp["features"][exact]
        nbad = int((ours != theirs).any(axis=1).sum())
        print(f"[descriptors] keypoints with bit-identical (x, y, angle): {int(exact.sum())}; descriptors differing: {nbad}")
        assert nbad == 0
        return missed

    assert not dm[:, 61:].any()
    # REPORT ONLY for the statistical bars: the whole-pipeline reference run merges sublevels with a live data race, its outcome
    # differs from run to run, and a bar it misses says nothing about this library.  The exact statements (set equality,
    # positions, descriptors against the race-free reference) are test_detector_vs_serialized_reference and the 1920x1088 test;
    # the descriptor comparison inside compare() is exact and stays asserted.
    missed = compare(ref_left["pts"])
    if missed:
        print(f"[keypoints] (report only) the racy reference run differs beyond the usual bars: {missed}")


def test_detector_vs_serialized_reference():
    """Extrema merge + radius NMS + refinement + orientation + M-LDB against the reference's own kernels run race-free
    (bindings.RefAkazer.detect_serialized): keypoint SETS must be equal, positions within 1e-4 px, descriptors of
    keypoints with identical (x, y, angle) bit-identical."""
    if not B.have_ref():
        pytest.skip("reference not built")
    for name, img in (("left.pgm", left_image()), ("right.pgm", right_image())):
        h, w = img.shape
        r = B.RefAkazer(w, h, w)
        rp, _, k = r.detect_serialized(dev(img)[0], max_pts=30000)
        r.close()
        ctx = ab().Context(w, h, max_batch=1, max_pts=30000, kcontrast_override=k)
        counts, kpts, desc = ctx.detect_and_compute(dev(img))
        ctx.sync()
        mine = _kp_array(counts, kpts)
        dm = desc[0].cpu().numpy()
        ctx.close()
        from scipy.spatial import cKDTree
        d, j = cKDTree(np.stack([mine["x"], mine["y"]], 1)).query(np.stack([rp["x"], rp["y"]], 1))
        same = (d <= 1e-4) & (mine["layer"][j] == rp["octave"])
        da = np.abs(mine["angle"][j] - rp["angle"])
        da = np.minimum(da, 2 * np.pi - da)
        exact = same & (mine["angle"][j].view(np.uint32) == rp["angle"].view(np.uint32)) & \
            (mine["x"][j].view(np.uint32) == rp["x"].view(np.uint32)) & (mine["y"][j].view(np.uint32) == rp["y"].view(np.uint32))
        nbad = int((dm[j[exact]][:, :61] != rp["features"][exact]).any(axis=1).sum())
        print(f"\n[{name} | serialised reference] ours={len(mine)} reference={len(rp)} same position+layer={same.mean():.5f} "
              f"max |dpos|={d.max():.2e} angle<=1e-4: {(da[same] <= 1e-4).mean():.5f} exact (x,y,angle)={int(exact.sum())} descriptors differing={nbad}")
        assert len(mine) == len(rp) and same.all()
        assert np.array_equal(mine["size"][j], rp["size"])
        assert (da <= 1e-4).mean() >= 0.995
        assert nbad == 0 and exact.mean() >= 0.75          # the rest differ in the last bits of the angle (App. B-4)


@needs_ref
def test_1080p_vs_reference_outside_the_uninitialised_band():
    """The benchmark size.  At 1920x1080 the reference's blur kernels read uninitialised shared-memory rows in the last tile
    row of octaves 0 and 2 (App. B-7); the garbage spreads upwards by a few pixels per level.  Everything above that band
    must be bit-identical: planes on the top 80 % of the rows, keypoints / descriptors with y < 0.75 H."""
    w, h = 1920, 1080
    img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=51))
    pitch = 1920
    r = B.RefAkazer(w, h, pitch)
    rp, planes, k = r.detect_serialized(dev(img)[0], max_pts=60000)
    r.close()
    ctx = ab().Context(w, h, max_batch=1, max_pts=60000, kcontrast_override=k)
    counts, kpts, desc = ctx.detect_and_compute(dev(img))
    ctx.sync()
    names = ["Lt", "det", "Lx", "Ly"]
    first_bad = {}
    # rows that must agree: the garbage enters at the bottom edge and climbs one row per diffusion step, so the band is
    # small at octaves 0-2 and grows to most of the 135-row plane over the 90 steps of octave 3
    top_frac = [0.8] * 12 + [0.75, 0.6, 0.42, 0.2]
    for l in range(ctx.num_levels):
        for which in range(4):
            a, b = bits(ctx.plane(l, which)), bits(planes[l][which])
            top = int(a.shape[0] * top_frac[l])
            assert np.array_equal(a[:top], b[:top]), f"level {l} {names[which]}: differs above the band"
            rows = np.nonzero((a != b).any(axis=1))[0]
            if len(rows):
                first_bad[(l, names[which])] = (int(rows[0]), a.shape[0])
    print(f"\n[1080p] first differing row (of rows) per plane, i.e. the reference's uninitialised band: "
          f"{[(k, v) for k, v in sorted(first_bad.items()) if k[1] == 'Lt']}")
    mine = _kp_array(counts, kpts)
    dm = desc[0].cpu().numpy()
    ctx.close()
    # Keypoints.  The uninitialised rows the reference reads are whatever the previous kernel left in shared memory: on a
    # fresh GPU they can be huge or NaN, and octave-3 candidates inside the band then suppress true keypoints of the finer
    # octaves around them through the radius NMS (seen once as a failure of a 99.5 % bar on y < 0.7 H in the first process
    # of a fresh box).  The set must be IDENTICAL where no level is contaminated and no contaminated candidate is within
    # NMS reach: level 15 is clean above row 27 of 135 (y < 216 at full resolution) and an octave-3 keypoint suppresses within
    # ~32 full-resolution pixels, so compare y < 0.15 H; the larger
    # region is reported, not asserted.
    def keyset(a, layer_field, ymax, lmax):
        sel = (a["y"] < ymax) & (a[layer_field] < lmax)
        return {(int(q[layer_field]), q["y"].view(np.uint32).item(), q["x"].view(np.uint32).item()) for q in a[sel]}
    ours, theirs = keyset(mine, "layer", 0.7 * h, 12), keyset(rp, "octave", 0.7 * h, 12)
    print(f"[1080p] keypoints (octaves 0-2, y < 0.7 H): ours={len(ours)} reference={len(theirs)} common={len(ours & theirs)}")
    ours, theirs = keyset(mine, "layer", 0.15 * h, 16), keyset(rp, "octave", 0.15 * h, 16)
    print(f"[1080p] keypoints (all octaves, y < 0.15 H): ours={len(ours)} reference={len(theirs)} common={len(ours & theirs)}")
    assert len(ours) > 30 and ours == theirs
    from scipy.spatial import cKDTree
    d, j = cKDTree(np.stack([mine["x"], mine["y"]], 1)).query(np.stack([rp["x"], rp["y"]], 1))
    exact = (rp["y"] < 0.55 * h) & (rp["octave"] < 8) & (d == 0) & (mine["angle"][j].view(np.uint32) == rp["angle"].view(np.uint32))
    nbad = int((dm[j[exact]][:, :61] != rp["features"][exact]).any(axis=1).sum())
    print(f"[1080p] descriptors compared (octaves 0-1, y < 0.55 H, identical x, y, angle): {int(exact.sum())}, differing: {nbad}")
    assert exact.sum() > 200 and nbad == 0


def test_orientation_and_descriptor_given_reference_keypoints(ref_left):
    """SURVEY App. F rows 'Orientation' and 'M-LDB': feed the reference's keypoints (and, for the descriptor, the
    reference's angle) through our kernels on bit-identical planes: descriptors must agree for 100 % of keypoints."""
    img = ref_left["img"]
    h, w = img.shape
    refp = ref_left["pts"]
    n = len(refp)
    ctx = ab().Context(w, h, max_batch=1, max_pts=max(n, 1), kcontrast_override=ref_left["k"])
    ctx.build_scale_space(dev(img))
    kp = np.zeros(n, dtype=ab().KEYPOINT_DTYPE)
    kp["x"], kp["y"], kp["size"], kp["angle"], kp["layer"] = refp["x"], refp["y"], refp["size"], refp["angle"], refp["octave"]
    kt = torch.from_numpy(kp.view(np.int32).reshape(1, n, 8)).cuda()
    ct = torch.tensor([n], dtype=torch.int32, device="cuda")
    dt = torch.zeros(1, n, 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.describe(ct, kt, dt)
    ctx.sync()
    ours = dt[0].cpu().numpy()
    bad = (ours[:, :61] != refp["features"]).any(axis=1)
    print(f"\n[M-LDB | reference keypoints+angle] {n} keypoints, {int(bad.sum())} descriptors differ")
    assert not bad.any()
    ctx.orient(ct, kt)
    ctx.sync()
    ang = ab().keypoints_from_words(kt[0].cpu().numpy())["angle"]
    da = np.abs(ang - refp["angle"])
    da = np.minimum(da, 2 * np.pi - da)
    print(f"[orientation | reference keypoints] max diff {da.max():.3e} rad; bit-identical {np.mean(ang.view(np.uint32) == refp['angle'].view(np.uint32)):.4f}")
    assert (da <= 1e-4).mean() >= 0.995
    ctx.close()


def _describe_given(ctx, refp):
    n = len(refp)
    kp = np.zeros(n, dtype=ab().KEYPOINT_DTYPE)
    kp["x"], kp["y"], kp["size"], kp["angle"], kp["layer"] = refp["x"], refp["y"], refp["size"], refp["angle"], refp["octave"]
    kt = torch.from_numpy(kp.view(np.int32).reshape(1, n, 8)).cuda()
    ct = torch.tensor([n], dtype=torch.int32, device="cuda")
    dt = torch.zeros(1, n, 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.describe(ct, kt, dt)
    ctx.sync()
    return dt[0].cpu().numpy()


def test_descriptor_kernels_agree_and_other_pattern_sizes_vs_reference(ref_left):
    """M-LDB has two kernels: k_describe_b (pattern size 10, the reference default) and the generic table-driven k_describe_s.
    Both must give the reference's bits on the reference's keypoints + angles; other pattern sizes (akaze.h:54 is a
    constructor argument) are checked against the reference built with the same size."""
    img = ref_left["img"]
    h, w = img.shape
    L = ab().lib()
    try:
        for generic in (0, 1):
            L.akz_set_describe_kernel(generic)
            ctx = ab().Context(w, h, max_batch=1, max_pts=len(ref_left["pts"]), kcontrast_override=ref_left["k"])
            ctx.build_scale_space(dev(img))
            ours = _describe_given(ctx, ref_left["pts"])
            ctx.close()
            assert np.array_equal(ours[:, :61], ref_left["pts"]["features"]), f"generic={generic}"
            assert not ours[:, 61:].any()
    finally:
        L.akz_set_describe_kernel(0)
    for pattern in (6, 13):
        r = B.RefAkazer(w, h, w, pattern=pattern)
        pts, planes, k = r.detect_keep(dev(img)[0], max_pts=30000)
        r.close()
        ctx = ab().Context(w, h, max_batch=1, max_pts=len(pts), kcontrast_override=k, descriptor_pattern_size=pattern)
        ctx.build_scale_space(dev(img))
        ours = _describe_given(ctx, pts)
        ctx.close()
        bad = (ours[:, :61] != pts["features"]).any(axis=1)
        # The reference does not clamp sample positions (akazed.cu:1917-1919): patches larger than the detector's border margin
        # (which is sized for pattern 10) read whatever lies beside the plane in its pyramid.  We clamp (DESIGN: deliberate
        # differences), so only keypoints whose rotated patch stays inside the level are comparable.
        o = pts["octave"] // 4
        ratio = 1.0 / (1 << o)
        reach = np.floor(pts["size"] + 0.5) * pattern * np.sqrt(2.0) + 2.0
        xf, yf = pts["x"] * ratio, pts["y"] * ratio
        inside = (xf - reach >= 0) & (yf - reach >= 0) & (xf + reach <= (w >> o) - 1) & (yf + reach <= (h >> o) - 1)
        print(f"\n[M-LDB | pattern {pattern}] {len(pts)} keypoints, {int(bad.sum())} descriptors differ, {int((bad & inside).sum())} of them "
              f"among the {int(inside.sum())} whose patch stays inside the level")
        assert inside.mean() > 0.8 and not (bad & inside).any(), pattern


def test_keypoints_equal_cpu_oracle():
    """Detector (extrema merge, NMS, refinement, raster order) is bit-identical to the CPU oracle."""
    for (w, h, seed) in [(640, 480, 5), (333, 250, 6)]:
        img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=seed))
        ctx = ab().Context(w, h, max_batch=1, max_pts=20000)
        counts, kpts, desc = ctx.detect_and_compute(dev(img))
        ctx.sync()
        mine = _kp_array(counts, kpts)
        P = B.OraclePyramid(w, h)
        P.build(img)
        ok = P.detect()
        assert len(mine) == len(ok) and len(ok) > 20, (len(mine), len(ok))
        for fld in ("ix", "iy", "layer"):
            assert np.array_equal(mine[fld], ok[fld]), fld
        for fld in ("x", "y", "size", "response"):
            assert np.array_equal(mine[fld].view(np.uint32), ok[fld].view(np.uint32)), fld
        ok = P.describe(ok)
        da = np.abs(mine["angle"] - ok["angle"])
        da = np.minimum(da, 2 * np.pi - da)
        assert (da <= 1e-4).mean() >= 0.99, da.max()
        P.close()
        ctx.close()


def test_five_octaves_equal_cpu_oracle():
    """configs[4] uses 5 octaves x 4 sublevels (FED cycles of up to 57 steps): planes of every level and the keypoints are
    bit-identical to the CPU oracle at a size whose fifth octave is 81x81."""
    w, h = 1296, 1300
    img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=31))
    ctx = ab().Context(w, h, noctaves=5, max_batch=1, max_pts=30000)
    assert ctx.num_levels == 20 and ctx.level_info(19)["nsteps"] == 57
    counts, kpts, desc = ctx.detect_and_compute(dev(img))
    ctx.sync()
    P = B.OraclePyramid(w, h, noctaves=5)
    P.build(img)
    assert P.levels == 20
    names = ["Lt", "det", "Lx", "Ly"]
    for l in (0, 3, 4, 11, 15, 16, 17, 18, 19):
        for which in range(4):
            assert_bits_equal(ctx.plane(l, which), P.plane(l, which), f"level {l} {names[which]}")
    mine = _kp_array(counts, kpts)
    ok = P.detect()
    assert len(mine) == len(ok) and len(ok) > 100
    for fld in ("ix", "iy", "layer"):
        assert np.array_equal(mine[fld], ok[fld]), fld
    for fld in ("x", "y", "response"):
        assert np.array_equal(mine[fld].view(np.uint32), ok[fld].view(np.uint32)), fld
    assert mine["layer"].max() >= 16                      # the fifth octave contributes keypoints
    P.close()
    ctx.close()


def test_u8_ingest_batching_host_api_and_determinism():
    w, h = 640, 480
    frames8 = np.stack([B.synth_shapes_u8(w, h, seed=s) for s in range(5)])
    framesf = B.u8_to_unit(frames8)
    ctx = ab().Context(w, h, max_batch=2, max_pts=8000)
    c1, k1, d1 = ctx.detect_and_compute(torch.from_numpy(framesf).cuda())        # 5 frames through chunks of 2
    ctx.sync()
    c2, k2, d2 = ctx.detect_and_compute(torch.from_numpy(frames8).cuda())        # u8 ingest
    ctx.sync()
    assert torch.equal(c1, c2) and int(c1.min()) > 20
    for f in range(5):
        n = int(c1[f])
        assert torch.equal(k1[f, :n], k2[f, :n]) and torch.equal(d1[f, :n], d2[f, :n])
    # one frame at a time == batched
    for f in (0, 3):
        cs, ks, ds = ctx.detect_and_compute(torch.from_numpy(framesf[f:f + 1]).cuda())
        ctx.sync()
        n = int(cs[0])
        assert n == int(c1[f]) and torch.equal(ks[0, :n], k1[f, :n]) and torch.equal(ds[0, :n], d1[f, :n])
    # host API == device API
    hc, hk, hd = ctx.detect_and_compute_host(frames8)
    assert np.array_equal(hc, c1.cpu().numpy())
    for f in range(5):
        n = int(hc[f])
        assert np.array_equal(hk[f, :n].view(np.int32).reshape(n, 8), k1[f, :n].cpu().numpy())
        assert np.array_equal(hd[f, :n], d1[f, :n].cpu().numpy())
    # determinism: ten runs, identical bits (App. F)
    for _ in range(10):
        c3, k3, d3 = ctx.detect_and_compute(torch.from_numpy(framesf).cuda())
        ctx.sync()
        assert torch.equal(c1, c3) and torch.equal(k1, k3) and torch.equal(d1, d3)
    ctx.close()


def test_sparse_detector_leaves_the_key_map_clean():
    """The detector's key map and its occupancy bitmap are cleared only where k_extrema wrote them (k_clear_map): a context that
    has seen keypoint-heavy frames, full chunks and partial chunks must return exactly what a fresh context returns."""
    w, h = 640, 480
    noise = B.u8_to_unit(np.stack([B.synth_noise_u8(w, h, seed=s) for s in range(4)]))
    shapes = B.u8_to_unit(np.stack([B.synth_shapes_u8(w, h, seed=70 + s) for s in range(3)]))
    fresh = ab().Context(w, h, max_batch=4, max_pts=20000)
    want = [t.clone() for t in fresh.detect_and_compute(torch.from_numpy(shapes).cuda())]
    fresh.sync()
    fresh.close()
    ctx = ab().Context(w, h, max_batch=4, max_pts=20000)
    cn, _, _ = ctx.detect_and_compute(torch.from_numpy(noise).cuda())              # a full chunk with many candidates everywhere
    ctx.sync()
    assert int(cn.min()) > int(want[0].max())
    ctx.detect_and_compute(torch.from_numpy(noise[:1]).cuda(), False)               # detect only, one frame
    got = ctx.detect_and_compute(torch.from_numpy(shapes).cuda())                   # a partial chunk
    ctx.sync()
    assert torch.equal(got[0], want[0])
    for f in range(3):
        n = int(want[0][f])
        assert torch.equal(got[1][f, :n], want[1][f, :n]) and torch.equal(got[2][f, :n], want[2][f, :n])
    ctx.close()


def test_two_lanes_equal_one_lane():
    """akz_options.lanes = 2 (chunks alternate between two streams with their own pyramids, AKZ_NSET staging buffers and
    result sets in the host pipeline) returns exactly what the single-lane context returns: device API, host API with
    float and u8 frames, the integer pipeline, and a batch that is not a multiple of the chunk size."""
    w, h = 640, 480
    nfr = 11
    frames8 = np.stack([B.synth_shapes_u8(w, h, seed=40 + s) for s in range(nfr)])
    framesf = B.u8_to_unit(frames8)
    one = ab().Context(w, h, max_batch=2, max_pts=6000, lanes=1)
    two = ab().Context(w, h, max_batch=2, max_pts=6000, lanes=2)
    dev = torch.from_numpy(framesf).cuda()
    c1, k1, d1 = one.detect_and_compute(dev)
    one.sync()
    for _ in range(3):
        c2, k2, d2 = two.detect_and_compute(dev)
        two.sync()
        assert torch.equal(c1, c2) and int(c1.min()) > 20
        for f in range(nfr):
            n = int(c1[f])
            assert torch.equal(k1[f, :n], k2[f, :n]) and torch.equal(d1[f, :n], d2[f, :n])
    for frames in (framesf, frames8):
        hc, hk, hd = two.detect_and_compute_host(frames)
        assert np.array_equal(hc, c1.cpu().numpy())
        for f in range(nfr):
            n = int(hc[f])
            assert np.array_equal(hk[f, :n].view(np.int32).reshape(n, 8), k1[f, :n].cpu().numpy())
            assert np.array_equal(hd[f, :n], d1[f, :n].cpu().numpy())
    dev8 = torch.from_numpy(frames8).cuda()
    f1 = one.fast_detect_and_compute(dev8)
    one.sync()
    f2 = two.fast_detect_and_compute(dev8)
    two.sync()
    assert torch.equal(f1[0], f2[0])
    for f in range(nfr):
        n = int(f1[0][f])
        assert torch.equal(f1[1][f, :n], f2[1][f, :n]) and torch.equal(f1[2][f, :n], f2[2][f, :n])
    hf = two.detect_and_compute_host(frames8, fast=True)
    assert np.array_equal(hf[0], f1[0].cpu().numpy())
    assert two.launches > 0
    one.close(); two.close()


def test_small_image_drops_octaves_and_empty_image():
    # 200x150: octave 1 is 100x75 (h < 80) -> one octave only (akaze.cpp:215-219, App. B-12)
    ctx = ab().Context(200, 150, max_batch=1, max_pts=1000)
    assert ctx.num_levels == 4
    z = torch.zeros(1, 150, 200, device="cuda")
    c, k, d = ctx.detect_and_compute(z)                   # a flat image has no keypoints: must not fail (App. B-9)
    ctx.sync()
    assert int(c[0]) == 0
    ctx.close()
    ctx = ab().Context(640, 480, max_batch=1, max_pts=50)   # overflow: count clamps to max_pts (akaze.cpp:451)
    img = B.u8_to_unit(B.synth_noise_u8(640, 480, seed=2))
    c, k, d = ctx.detect_and_compute(dev(img))
    ctx.sync()
    assert int(c[0]) == 50
    ctx.close()


def test_full_size_1080p_properties():
    """BASELINE configs[2] size (1920x1080, 4x4): size-independent properties instead of an oracle pass --
    the fused production kernels equal the one-kernel-per-stage path bit for bit on every plane, and a frame's result
    does not depend on its position in the batch or on the chunking."""
    w, h = 1920, 1080
    f0 = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=41))
    f1 = B.u8_to_unit(B.synth_noise_u8(w, h, seed=42))
    t = torch.from_numpy(np.stack([f0, f1, f0])).cuda()
    planes = {}
    for fused in (0, 1):
        ctx = ab().Context(w, h, fused=fused, max_batch=3, max_pts=40000)
        ctx.build_scale_space(t)
        ctx.sync()
        planes[fused] = [[ctx.plane(l, which, 1) for which in range(4)] for l in (0, 1, 5, 10, 15)]
        if fused:
            c3, k3, d3 = ctx.detect_and_compute(t)
            ctx.sync()
        ctx.close()
    for a, b in zip(planes[0], planes[1]):
        for which in range(4):
            assert_bits_equal(a[which], b[which], "1080p stages vs fused")
    n0, n1 = int(c3[0]), int(c3[1])
    assert n0 == int(c3[2]) and n0 > 500 and n1 > 5000
    assert torch.equal(k3[0, :n0], k3[2, :n0]) and torch.equal(d3[0, :n0], d3[2, :n0])          # position in the batch
    ctx = ab().Context(w, h, max_batch=2, max_pts=40000)                                              # different chunking
    c2, k2, d2 = ctx.detect_and_compute(t)
    ctx.sync()
    assert torch.equal(c2, c3)
    for f in range(3):
        n = int(c3[f])
        assert torch.equal(k2[f, :n], k3[f, :n]) and torch.equal(d2[f, :n], d3[f, :n])
    ctx.close()


def test_keypoint_capacity_is_clamped_in_raster_order():
    """App. B-10: more survivors than max_pts -> the count is clamped and the FIRST max_pts keypoints in raster order are
    kept (the reference keeps an order-dependent subset and indexes past the buffer)."""
    w, h = 640, 480
    img = B.u8_to_unit(B.synth_noise_u8(w, h, seed=3))
    big = ab().Context(w, h, max_batch=1, max_pts=20000)
    cb, kb, db = big.detect_and_compute(dev(img))
    big.sync()
    n = int(cb[0])
    assert n > 300
    small = ab().Context(w, h, max_batch=1, max_pts=200)
    cs, ks, ds = small.detect_and_compute(dev(img))
    small.sync()
    assert int(cs[0]) == 200
    assert torch.equal(ks[0, :200], kb[0, :200]) and torch.equal(ds[0, :200], db[0, :200])
    big.close(); small.close()


def test_report_overlap_with_opencv_akaze():
    """Report only (App. B-11: the reference differs from OpenCV/upstream AKAZE by design): fraction of our keypoints on
    left.pgm that have an OpenCV cv::AKAZE keypoint within 3 px, and the other way round."""
    cv2 = pytest.importorskip("cv2")
    img = left_image()
    h, w = img.shape
    ctx = ab().Context(w, h, max_batch=1, max_pts=30000)
    counts, kpts, _ = ctx.detect_and_compute(dev(img))
    ctx.sync()
    mine = _kp_array(counts, kpts)
    ctx.close()
    kp = cv2.AKAZE_create().detect(np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8), None)
    cvxy = np.array([k.pt for k in kp], dtype=np.float32)
    from scipy.spatial import cKDTree
    d1, _ = cKDTree(cvxy).query(np.stack([mine["x"], mine["y"]], 1))
    d2, _ = cKDTree(np.stack([mine["x"], mine["y"]], 1)).query(cvxy)
    print(f"\n[vs OpenCV cv::AKAZE on left.pgm] ours={len(mine)} opencv={len(kp)}; ours within 3 px of an OpenCV keypoint: {(d1 <= 3).mean():.3f}; "
          f"OpenCV within 3 px of ours: {(d2 <= 3).mean():.3f}")
    assert len(mine) > 1000 and len(kp) > 1000 and (d2 <= 3).mean() > 0.3


# ---------------------------------------------------------------------------------------------------------------
# matcher
# ---------------------------------------------------------------------------------------------------------------
def _planted(nq, nt, seed):
    q = B.random_descriptors(nq, seed)
    t = B.random_descriptors(nt, seed + 1)
    rng = np.random.default_rng(seed + 2)
    # plant near-duplicates, exact duplicates in the same and in different strides, and ties
    for i in range(0, min(nq, nt) // 4):
        j = int(rng.integers(0, nt))
        t[j] = q[i]
        flips = rng.integers(0, 486, size=int(rng.integers(0, 40)))
        for b in flips:
            t[j, b // 8] ^= np.uint8(1 << (b % 8))
        if i % 7 == 0 and j + 16 < nt:
            t[j + 16] = t[j]            # tie inside one stride: still accepted
        if i % 11 == 0 and j + 5 < nt:
            t[j + 5] = t[j]             # tie across strides: rejected
    t[:, 61:] = 0
    t[:, 60] &= 0x3F
    return q, t


@pytest.mark.parametrize("nq,nt", [(1000, 1500), (257, 16), (33, 4099), (1, 100)])
def test_matcher_vs_cpu_oracle(nq, nt):
    q, t = _planted(nq, nt, seed=nq + nt)
    ctx = ab().Context(0, 0)
    qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    torch.cuda.synchronize()
    r = ctx.match(qt, tt, ab().MATCH_COMPAT)
    ctx.sync()
    r = r.cpu().numpy()
    o = B.oracle_match(q, t, "compat")
    assert np.array_equal(r[:, :2], o), np.argwhere(r[:, :2] != o)[:5]
    r2 = ctx.match(qt, tt, ab().MATCH_KNN2)
    ctx.sync()
    o2 = B.oracle_match(q, t, "knn2")
    assert np.array_equal(r2.cpu().numpy(), o2)
    assert np.array_equal(ctx.match_host(q, t, ab().MATCH_KNN2), o2)
    # UNIQUE2 = the reference's gMatch acceptance (akazed.cu:2103) on top of the exact top-2
    r3 = ctx.match(qt, tt, ab().MATCH_UNIQUE2)
    ctx.sync()
    r3 = r3.cpu().numpy()
    ok = (o2[:, 0] >= 0) & (o2[:, 1] < 96) & ((o2[:, 2] < 0) | (o2[:, 1] < o2[:, 3]))
    assert np.array_equal(r3[:, 0], np.where(ok, o2[:, 0], -1)) and np.array_equal(r3[:, 1], np.where(ok, o2[:, 1], -1))
    assert np.array_equal(r3[:, 2:], o2[:, 2:])
    # sharded: split the train set in 3 ranges, merge the partial results (what the NCCL path does after its gather)
    for mode, ref_out in ((ab().MATCH_COMPAT, o), (ab().MATCH_KNN2, o2)):
        cuts = [0, nt // 3, (2 * nt) // 3, nt]
        parts = torch.stack([ctx.match(qt, tt[cuts[i]:cuts[i + 1]].contiguous(), mode, t_index_base=cuts[i], finalize=False) for i in range(3)])
        ctx.sync()
        m = ctx.match_merge(parts, mode, finalize=True)
        ctx.sync()
        m = m.cpu().numpy()
        assert np.array_equal(m[:, :ref_out.shape[1]], ref_out)
    ctx.close()


@pytest.mark.parametrize("nq,nt", [(1000, 1500), (300, 129), (2049, 4097), (5, 77)])
def test_matcher_tensor_core_kernel_equals_popc_kernel(nq, nt):
    """k_match_tc5 (tcgen05.mma kind::i8, accumulators in tensor memory), k_match_mma (mma.sync IMMA) and k_match (LOP3/POPC)
    give identical results in every mode, including ragged tiles, planted ties and a sharded train set."""
    q, t = _planted(nq, nt, seed=3 * nq + nt)
    ctx = ab().Context(0, 0)
    qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    L = ab().lib()
    try:
        for mode in (ab().MATCH_COMPAT, ab().MATCH_KNN2, ab().MATCH_UNIQUE2):
            out = {}
            for kern in (1, 2, 3):
                L.akz_set_match_kernel(kern)
                r = ctx.match(qt, tt, mode)
                cut = nt // 2
                parts = torch.stack([ctx.match(qt, tt[:cut].contiguous(), mode, t_index_base=0, finalize=False),
                                     ctx.match(qt, tt[cut:].contiguous(), mode, t_index_base=cut, finalize=False)])
                ctx.sync()
                m = ctx.match_merge(parts, mode, finalize=True)
                ctx.sync()
                out[kern] = (r.cpu().numpy(), m.cpu().numpy())
            for kern in (2, 3):
                assert np.array_equal(out[1][0], out[kern][0]), (mode, kern, np.argwhere(out[1][0] != out[kern][0])[:4])
                assert np.array_equal(out[1][1], out[kern][1])
            assert np.array_equal(out[1][0], out[1][1])
    finally:
        L.akz_set_match_kernel(0)
    ctx.close()


@pytest.mark.parametrize("nq,nt", [(1, 100), (257, 16), (33, 4099), (3000, 5000), (700, 1024), (256, 1025)])
def test_tcgen05_matcher_vs_cpu_oracle(nq, nt):
    """k_match_tc5 forced (kernel 3 = filter by range length, 4 = chunk filter on, 5 = off) against the CPU oracle in both
    modes: single query, fewer train descriptors than a tile, ragged last tiles, planted ties inside and across the
    16-strides, ranges that are exact multiples of the 1024-descriptor unit, and a sharded train set whose shard bases are
    not multiples of 16 (the class of a train index is (base + index) mod 16)."""
    q, t = _planted(nq, nt, seed=7 * nq + nt)
    o = B.oracle_match(q, t, "compat")
    o2 = B.oracle_match(q, t, "knn2")
    ctx = ab().Context(0, 0)
    qt, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    L = ab().lib()
    try:
        for kern in (3, 4, 5):
            L.akz_set_match_kernel(kern)
            r = ctx.match(qt, tt, ab().MATCH_COMPAT)
            ctx.sync()
            assert np.array_equal(r.cpu().numpy()[:, :2], o), (kern, np.argwhere(r.cpu().numpy()[:, :2] != o)[:5])
            r2 = ctx.match(qt, tt, ab().MATCH_KNN2)
            ctx.sync()
            assert np.array_equal(r2.cpu().numpy(), o2), (kern, np.argwhere(r2.cpu().numpy() != o2)[:5])
            if nt >= 40:
                cuts = [0, nt // 3 + 5, (2 * nt) // 3 + 3, nt]
                for mode, ref_out in ((ab().MATCH_COMPAT, o), (ab().MATCH_KNN2, o2)):
                    parts = torch.stack([ctx.match(qt, tt[cuts[i]:cuts[i + 1]].contiguous(), mode, t_index_base=cuts[i], finalize=False)
                                         for i in range(3)])
                    ctx.sync()
                    m = ctx.match_merge(parts, mode, finalize=True)
                    ctx.sync()
                    assert np.array_equal(m.cpu().numpy()[:, :ref_out.shape[1]], ref_out), (kern, mode)
    finally:
        L.akz_set_match_kernel(0)
    ctx.close()


def test_tcgen05_matcher_long_range_is_split():
    """A train set longer than the 2^19-descriptor range one block can index is split by akz_match; results equal the
    LOP3/POPC kernel's (which has no such limit below 2^22)."""
    nq, nt = 300, 600_000
    q = B.random_descriptors(nq, 11)
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    tt = torch.randint(0, 256, (nt, 64), dtype=torch.uint8, device="cuda", generator=g)
    tt[:, 61:] = 0
    tt[:, 60] &= 0x3F
    qt = torch.from_numpy(q).cuda()
    tt[599_999] = qt[7]                         # best match in the very last row
    tt[524_288] = qt[8]                         # and at the first index beyond one block's range
    ctx = ab().Context(0, 0)
    L = ab().lib()
    try:
        out = {}
        for kern in (1, 3):
            L.akz_set_match_kernel(kern)
            out[kern] = (ctx.match(qt, tt, ab().MATCH_KNN2).cpu().numpy(), ctx.match(qt, tt, ab().MATCH_COMPAT).cpu().numpy())
        assert np.array_equal(out[1][0], out[3][0]) and np.array_equal(out[1][1], out[3][1])
        assert out[3][0][7, 0] == 599_999 and out[3][0][7, 1] == 0 and out[3][0][8, 0] == 524_288
    finally:
        L.akz_set_match_kernel(0)
    ctx.close()


@needs_ref
def test_matcher_vs_reference():
    # nt must be a multiple of 16: gHammingMatch calls __syncthreads() inside a loop whose trip count differs per
    # thread otherwise (akazed.cu:2176-2187, App. B-15) -- on sm_100a that deadlocks (observed: the kernel never returns)
    nq, nt = 1200, 1696
    q, t = _planted(nq, nt, seed=99)
    pq = np.zeros(nq, dtype=B.REF_POINT)
    pt = np.zeros(nt, dtype=B.REF_POINT)
    pq["features"], pt["features"] = q[:, :61], t[:, :61]
    pt["x"], pt["y"] = np.arange(nt), np.arange(nt) * 2
    dq = torch.from_numpy(pq.view(np.uint8).reshape(-1)).cuda()
    dt = torch.from_numpy(pt.view(np.uint8).reshape(-1)).cuda()
    torch.cuda.synchronize()
    B.ref().ref_hMatch(C.c_void_p(dq.data_ptr()), nq, C.c_void_p(dt.data_ptr()), nt)
    torch.cuda.synchronize()
    rq = dq.cpu().numpy().view(B.REF_POINT)
    ctx = ab().Context(0, 0)
    r = ctx.match(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), ab().MATCH_COMPAT)
    ctx.sync()
    r = r.cpu().numpy()
    # the reference adds the popcount of 3 uninitialised shared-memory bytes to every distance of a query (App. B-6):
    # indices must agree wherever that constant cannot flip the <96 gate
    agree = (r[:, 0] == rq["match"])
    gate = (r[:, 0] >= 0) & (r[:, 1] >= 96 - 24) & (rq["match"] < 0)       # only the <96 gate may flip, and only this way
    print(f"\n[matcher vs reference] {nq}x{nt}: index agreement {agree.mean():.5f}; explained by the gate: {int((~agree & gate).sum())}; "
          f"unexplained: {int((~agree & ~gate).sum())}")
    assert (~agree & ~gate).sum() == 0
    assert agree.mean() >= 0.97
    both = (r[:, 0] >= 0) & (rq["match"] >= 0)
    off = rq["distance"][both] - r[both, 1]
    assert off.min() >= 0 and off.max() <= 24
    ctx.close()


def test_matcher_full_size_properties():
    """BASELINE metric 2 shape (10k x 10k): size-independent properties instead of an oracle pass."""
    n = 10000
    q = B.random_descriptors(n, 1)
    t = q[np.random.default_rng(5).permutation(n)].copy()
    perm_inv = np.argsort(np.random.default_rng(5).permutation(n))
    ctx = ab().Context(0, 0)
    r = ctx.match(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), ab().MATCH_KNN2)
    ctx.sync()
    r = r.cpu().numpy()
    assert np.array_equal(r[:, 0], perm_inv) and not r[:, 1].any()          # every query finds its own copy at distance 0
    assert (r[:, 3] > 150).all()                                           # second best of random 486-bit strings
    ctx.close()


def test_match_pairs_equals_per_pair_matching():
    """akz_match_pairs (batched consecutive-frame matcher of the 4K stream, counts read on the device) against akz_match pair by
    pair, all modes; frames with zero, few and many keypoints."""
    rng = np.random.default_rng(11)
    mp, nf = 3000, 6
    counts = np.array([1500, 2999, 0, 17, 3000, 700], dtype=np.int32)
    desc = np.zeros((nf, mp, 64), dtype=np.uint8)
    base = B.random_descriptors(3000, 3)
    for f in range(nf):
        d = base[rng.permutation(3000)].copy()
        flip = rng.random((3000, 64)) < 0.02
        d ^= (flip * rng.integers(0, 256, (3000, 64))).astype(np.uint8)
        d[:, 61:] = 0
        d[:, 60] &= 0x3F
        desc[f, :counts[f]] = d[:counts[f]]
    ctx = ab().Context(0, 0, max_pts=mp)
    dd, dc = torch.from_numpy(desc).cuda(), torch.from_numpy(counts).cuda()
    for mode in (ab().MATCH_COMPAT, ab().MATCH_KNN2, ab().MATCH_UNIQUE2):
        out = torch.full((nf, mp, 4), -7, dtype=torch.int32, device="cuda")
        ctx.match_pairs(dd, dc, mode, out=out)
        ctx.sync()
        out = out.cpu().numpy()
        for f in range(1, nf):
            if counts[f] == 0:
                continue
            r = ctx.match(dd[f, :counts[f]], dd[f - 1, :counts[f - 1]], mode)
            ctx.sync()
            r = r.cpu().numpy()
            cols = slice(0, 2) if mode == ab().MATCH_COMPAT else slice(0, 4)
            assert np.array_equal(out[f, :counts[f], cols], r[:, cols]), (mode, f)
            assert (out[f, counts[f]:] == -7).all()                    # rows beyond the count are not touched
    ctx.close()


def test_match_sharded_with_a_one_rank_communicator():
    """akz_match_sharded end to end on one GPU: NCCL communicator of one rank created through akz_comm_unique_id / akz_comm_init,
    akz_match(finalize = 0) -> ncclAllGather on the context's stream -> akz_match_merge.  Must equal akz_match.  (World sizes
    2..8 run in bench.py under torchrun, which asserts the same equality against the unsharded result.)"""
    q, t = _planted(3000, 20000, seed=5)
    ctx = ab().Context(0, 0)
    try:
        ctx.comm_init(1, 0, ab().comm_unique_id())
    except ab().AkazeError as e:
        ctx.close()
        pytest.skip(f"NCCL not usable here: {e}")
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    for mode in (ab().MATCH_KNN2, ab().MATCH_COMPAT):
        a = ctx.match_sharded(dq, dt, 0, mode)
        b_ = ctx.match(dq, dt, mode)
        ctx.sync()
        cols = 2 if mode == ab().MATCH_COMPAT else 4
        assert torch.equal(a[:, :cols], b_[:, :cols])
        # a train range that starts at a global offset keeps global indices
        c = ctx.match_sharded(dq, dt[1024:], 1024, mode)
        d = ctx.match(dq, dt[1024:], mode, t_index_base=1024)
        ctx.sync()
        assert torch.equal(c[:, :cols], d[:, :cols])
    ctx.comm_destroy()
    ctx.close()


@pytest.mark.parametrize("fast", [False, True])
def test_small_batch_graph_path_equals_the_batched_path(fast):
    """Contexts of batch <= 4 run the octave chains on their own streams and replay a chunk as a CUDA graph from the second
    call with the same buffers on (the path behind akaze::Akazer).  Every call -- eager, captured, replayed, and after the
    input buffer's CONTENT changed -- must give exactly what a large-batch context (one stream, no graph) gives."""
    w, h = 1280, 960
    imgs = [B.read_pgm(os.path.join(B.REF_DATA, n)) if os.path.exists(os.path.join(B.REF_DATA, n)) else B.synth_shapes_u8(w, h, seed=7 + i)
            for i, n in enumerate(("left.pgm", "right.pgm"))]
    conv = (lambda a: a) if fast else B.u8_to_unit
    big = ab().Context(w, h, max_batch=8, max_pts=8000)
    run_big = big.fast_detect_and_compute if fast else big.detect_and_compute
    expect = []
    for im in imgs:
        c, k, d = run_big(torch.from_numpy(conv(im)[None]).cuda())
        big.sync()
        n = int(c[0].cpu())
        expect.append((n, k[0, :n].cpu().numpy().copy(), d[0, :n].cpu().numpy().copy()))
    big.close()
    for nb in (1, 2):
        small = ab().Context(w, h, max_batch=nb, max_pts=8000)
        run_small = small.fast_detect_and_compute if fast else small.detect_and_compute
        buf = torch.from_numpy(conv(imgs[0])[None]).cuda()
        out = small.alloc_results(1, True)
        l0 = small.launches
        for it in range(5):
            src = imgs[it % 2]
            buf.copy_(torch.from_numpy(conv(src)[None]).cuda())
            torch.cuda.synchronize()
            c, k, d = run_small(buf, out=out)
            small.sync()
            n, ek, ed = expect[it % 2]
            assert int(c[0].cpu()) == n, (nb, it)
            assert np.array_equal(k[0, :n].cpu().numpy(), ek) and np.array_equal(d[0, :n].cpu().numpy(), ed), (nb, it)
        assert small.launches - l0 > 5 * 40                              # replayed launches are counted too
        small.close()


def _full_parity_vs_serialised_reference(img, noct, tag, clean_level, kp_region, min_common=1.0):
    """Whole pipeline against the reference run race-free (bindings.RefAkazer.detect_serialized) on one frame: every plane of
    every level that the reference computes without reading uninitialised memory (clean_level(l) -> fraction of rows from the
    top that must agree, 1.0 = all), the keypoint SET inside kp_region(points) -> mask, positions <= 1e-4 px, and the descriptors
    of all keypoints with bit-identical (x, y, angle)."""
    from scipy.spatial import cKDTree
    h, w = img.shape
    r = B.RefAkazer(w, h, w, noctaves=noct)
    rp, planes, k = r.detect_serialized(dev(img)[0], max_pts=60000)
    r.close()
    ctx = ab().Context(w, h, noctaves=noct, max_batch=1, max_pts=60000, kcontrast_override=k)
    counts, kpts, desc = ctx.detect_and_compute(dev(img))
    ctx.sync()
    assert ctx.num_levels == len(planes) == 4 * noct
    names = ["Lt", "det", "Lx", "Ly"]
    nfull = 0
    for l in range(ctx.num_levels):
        frac = clean_level(l)
        for which in range(4):
            a, b = bits(ctx.plane(l, which)), bits(planes[l][which])
            top = int(round(a.shape[0] * frac))
            if frac >= 1.0:
                assert_bits_equal(ctx.plane(l, which), planes[l][which], f"{tag} level {l} {names[which]}")
                nfull += 1
            else:
                assert np.array_equal(a[:top], b[:top]), f"{tag} level {l} {names[which]}: differs above the reference's band"
    mine = _kp_array(counts, kpts)
    dm = desc[0].cpu().numpy()
    ctx.close()
    mm, rm = kp_region(mine, "layer"), kp_region(rp, "octave")
    key = lambda a, f, m: {(int(q[f]), q["y"].view(np.uint32).item(), q["x"].view(np.uint32).item()) for q in a[m]}
    ours, theirs = key(mine, "layer", mm), key(rp, "octave", rm)
    d, j = cKDTree(np.stack([mine["x"], mine["y"]], 1)).query(np.stack([rp["x"][rm], rp["y"][rm]], 1))
    rr = rp[rm]
    exact = (d == 0) & (mine["layer"][j] == rr["octave"]) & (mine["angle"][j].view(np.uint32) == rr["angle"].view(np.uint32))
    nbad = int((dm[j[exact]][:, :61] != rr["features"][exact]).any(axis=1).sum())
    da = np.abs(mine["angle"][j] - rr["angle"])
    da = np.minimum(da, 2 * np.pi - da)
    print(f"\n[{tag}] planes compared in full: {nfull} of {4 * len(planes)}; keypoints ours={len(ours)} reference={len(theirs)} common={len(ours & theirs)}; "
          f"angles within 1e-4 rad: {(da <= 1e-4).mean():.5f}; descriptors compared (identical x, y, angle): {int(exact.sum())}, differing: {nbad}")
    assert len(theirs) > 200
    if min_common >= 1.0:
        assert ours == theirs
    else:
        assert len(ours & theirs) >= min_common * max(len(ours), len(theirs))
    assert (da[d == 0] <= 1e-4).mean() >= 0.995
    assert nbad == 0 and exact.mean() >= 0.7
    return len(theirs)


@needs_ref
@pytest.mark.parametrize("content", ["shapes", "noise"])
def test_benchmark_scale_1920x1088_vs_reference_complete(content):
    """The benchmark-scale frame at the nearest size where the reference reads no uninitialised rows (1088 / 544 / 272 / 136 rows,
    SURVEY App. B-7): ALL 64 planes bit for bit, the COMPLETE keypoint set and every comparable descriptor, no band excluded.
    Both synthetic inputs of bench.py (shapes ~2 k keypoints, noise ~20 k)."""
    w, h = 1920, 1088
    gen = B.synth_noise_u8 if content == "noise" else B.synth_shapes_u8
    img = B.u8_to_unit(gen(w, h, seed=51))
    n = _full_parity_vs_serialised_reference(img, 4, f"1920x1088 {content}", lambda l: 1.0, lambda a, f: np.ones(len(a), dtype=bool))
    assert n > (10000 if content == "noise" else 1000)


@needs_ref
def test_4k_five_octaves_vs_reference():
    """BASELINE configs[4] geometry: 3840x2160, 5 octaves x 4 sublevels.  The reference reads uninitialised rows at octave 3
    (270 rows, App. B-7) and octave 4 inherits them through the octave transition: octaves 0-2 (48 planes) are compared in
    full, octaves 3 and 4 above the band the garbage reaches (it climbs one row per diffusion step from the bottom edge; the
    last level of octave 4, 57 steps on 135 rows, is left out).  Keypoints of octaves 0-2 in the upper 55 % of the frame: a
    contaminated coarse candidate can still suppress a true fine keypoint near it through the radius NMS, so the sets must
    agree to 99.5 % there rather than exactly; descriptors of all keypoints with identical (x, y, angle) are exact."""
    w, h = 3840, 2160
    img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=21, nshapes=600))
    top = {12: 0.75, 13: 0.6, 14: 0.42, 15: 0.2, 16: 0.7, 17: 0.45, 18: 0.15, 19: 0.0}
    _full_parity_vs_serialised_reference(img, 5, "3840x2160 5x4", lambda l: top.get(l, 1.0),
                                         lambda a, f: (a["y"] < 0.55 * h) & (a[f] < 12), min_common=0.995)


def test_fed_tma_variant_is_exact():
    """The measured alternative of k_fed4 (AKZ_FED_TMA=1: rows by cp.async.bulk + mbarrier instead of cp.async, see
    profiles/r02_fed4_tma_ab.txt) must give the same bits.  The knob is read once per process: run it in a child process."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "cuda-akaze_b200"); sys.path.insert(0, "tests")
import akaze_b200 as ab
for (w, h, n) in ((1920, 200, 3), (480, 270, 4), (250 * 4, 97, 7)):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    L = torch.rand(3, h, w, device="cuda", generator=g); G = torch.rand(3, h, w, device="cuda", generator=g)
    tau = (np.random.default_rng(n).random(n) * 0.2 + 0.01).astype(np.float32)
    outs = []
    for fused in (1, 0):
        c = ab.Context(0, 0, fused=fused, max_batch=3)
        d, t = torch.zeros_like(L), torch.zeros_like(L)
        c.fed_cycle(L, G, d, t, w, tau); c.sync(); outs.append(d); c.close()
    assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32)), (w, h, n)
print("tma-ok")
'''
    env = dict(os.environ, AKZ_FED_TMA="1", AKZ_FED_MIN_UNITS="0")
    out = subprocess.run([sys.executable, "-c", code], cwd=B.ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "tma-ok" in out.stdout, out.stderr[-1500:]


_LEVEL4_CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, "cuda-akaze_b200"); sys.path.insert(0, "tests")
import akaze_b200 as ab, bindings as B
total = 0
for (w, h, n, seed) in ((1920, 1088, 2, 3), (640, 480, 3, 5), (336, 250, 2, 7), (1000, 97, 4, 9)):
    imgs8 = np.stack([B.synth_noise_u8(w, h, seed=seed + i) if i % 2 else B.synth_shapes_u8(w, h, seed=seed + i) for i in range(n)])
    imgs = torch.from_numpy(np.stack([B.u8_to_unit(x) for x in imgs8])).cuda()
    imgs8 = torch.from_numpy(imgs8).cuda()
    for fast in (False, True):
        res = {}
        for fused in (1, 0):
            c = ab.Context(w, h, fused=fused, max_batch=n, max_pts=30000)
            l0 = c.launches
            if fast:
                cnt, kp, de = c.fast_detect_and_compute(imgs8)
            else:
                cnt, kp, de = c.detect_and_compute(imgs)
            c.sync()
            if fused: total += c.launches - l0
            get = c.plane_int if fast else c.plane
            planes = [[np.ascontiguousarray(get(l, which, f)).view(np.uint32).copy() for which in range(4)] for f in range(n) for l in range(c.num_levels)]
            res[fused] = (planes, cnt.cpu().numpy().copy(), kp.cpu().numpy().copy(), de.cpu().numpy().copy())
            c.close()
        for i, (a, b) in enumerate(zip(res[1][0], res[0][0])):
            for which in range(4):
                assert np.array_equal(a[which], b[which]), (w, h, fast, "frame/level", i, "plane", which, int((a[which] != b[which]).sum()))
        assert np.array_equal(res[1][1], res[0][1]), (w, h, fast, res[1][1], res[0][1])
        for f in range(n):
            m = int(res[1][1][f])
            assert m > 20, (w, h, fast, f, m)
            assert np.array_equal(res[1][2][f, :m], res[0][2][f, :m]) and np.array_equal(res[1][3][f, :m], res[0][3][f, :m]), (w, h, fast, f)
print("level4-ok", total)
'''


def test_level4_fused_level_kernel_is_exact():
    """k_level4 (level_stream.cu: blur + conductance + Lx / Ly / det of a level in one streaming kernel) against the per-stage
    kernels (fused = 0): every plane of every level bit for bit, the same keypoints and descriptors, float and integer pipeline,
    sizes with partial bands and partial strips.  AKZ_LEVEL4=2 forces the kernel at batch sizes where the size heuristic would not
    pick it (the knob is read once per process: child processes); with AKZ_LEVEL4=0 the same run needs more launches."""
    import subprocess
    import sys
    launches = {}
    for knob in ("2", "0"):
        env = dict(os.environ, AKZ_LEVEL4=knob)
        out = subprocess.run([sys.executable, "-c", _LEVEL4_CHILD], cwd=B.ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "level4-ok" in out.stdout, (knob, out.stdout[-500:], out.stderr[-1500:])
        launches[knob] = int(out.stdout.split("level4-ok")[1].split()[0])
    assert launches["2"] < launches["0"], launches
