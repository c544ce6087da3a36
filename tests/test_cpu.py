"""CPU-side tests (no GPU): the oracle against golden constants and fixtures, host logic of the product library
(FED schedule, taps, comparison table, C-ABI surface), the drop-in headers, and the multi-rank host logic."""
import ctypes as C
import hashlib
import os
import sys
import re
import subprocess

import numpy as np
import pytest

import bindings as B

ROOT = B.ROOT


def ab():
    import akaze_b200
    return akaze_b200


def fnv1a64(data: bytes):
    h = 0xcbf29ce484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


# ---- C ABI surface ------------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "akaze_b200.h")).read()
    declared = set(re.findall(r"AKZ_API\s+[\w\s\*]+?\b(akz_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    L = ab().lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(ab().EXPORTS)
    assert L.akz_version() >= 100


def test_communicator_id_needs_no_gpu():
    """akz_comm_unique_id (NCCL loaded at run time, no link-time dependency): 128 bytes, different on every call."""
    try:
        a, b = ab().comm_unique_id(), ab().comm_unique_id()
    except ab().AkazeError as e:
        pytest.skip(f"NCCL not available: {e}")
    assert len(a) == 128 and a != b and any(a)
    out = subprocess.check_output(["ldd", ab().LIB_PATH]).decode()
    assert "nccl" not in out


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ab().AkazeError):
        ab().Context(640, 480)


def test_option_validation_and_error_text():
    """akz_create rejects bad options with AKZ_E_INVALID before it touches the device; akz_last_error carries the reason.
    The reference has no such checks: it print-and-exits from inside CUDA calls (cuda_utils.h:18-37)."""
    L = ab().lib()
    for kw, needle in ((dict(width=640, height=480, max_scale=6), "octaves/sublevels"), (dict(width=640, height=480, noctaves=0), "octaves/sublevels"),
                       (dict(width=8, height=8), "frame size"), (dict(width=640, height=480, dthreshold=-1.0), "dthreshold"),
                       (dict(width=640, height=480, max_pts=0), "max_pts")):
        o = ab().default_options(**kw)
        h = C.c_void_p()
        rc = L.akz_create(C.byref(o), C.byref(h))
        assert rc == -1 and needle in L.akz_last_error().decode(), (kw, rc, L.akz_last_error())
    assert L.akz_create(None, None) == -1
    d = ab().default_options()
    assert (d.noctaves, d.max_scale, d.reordering, d.diffusivity, d.descriptor_pattern_size, d.max_pts) == (4, 4, 1, 1, 10, 10000)   # akaze.h:34-54, main.cpp:157
    assert abs(d.per - 0.7) < 1e-7 and abs(d.soffset - 1.6) < 1e-7 and abs(d.derivative_factor - 1.5) < 1e-7 and abs(d.dthreshold - 0.001) < 1e-9


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under cuda-akaze_b200/ may reference it."""
    for base, _, files in os.walk(os.path.join(ROOT, "cuda-akaze_b200")):
        if os.path.basename(base) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".so", ".o", ".log", ".pyc")):
                continue
            txt = open(os.path.join(base, f), errors="ignore").read()
            assert "oracle" not in txt.lower().replace("oracle-comparison", ""), os.path.join(base, f)
    out = subprocess.check_output(["ldd", ab().LIB_PATH]).decode()
    assert "oracle" not in out and "libref" not in out


def test_dropin_cxx_surface_is_exported_and_links():
    """The library is built with hidden visibility: every function the drop-in headers declare (akaze.h, akazed.h, fed.h:
    the names main.cpp / akaze.cpp call) must still be exported, and the C window used by the GPU tests must link and load
    (loading needs no GPU; no compute call is made here)."""
    out = subprocess.check_output(["nm", "-DC", "--defined-only", ab().LIB_PATH]).decode()
    for sym in ("akaze::initAkazeData(", "akaze::freeAkazeData(", "akaze::cuMatch(", "akaze::Akazer::Akazer()", "akaze::Akazer::~Akazer()",
                "akaze::Akazer::init(", "akaze::Akazer::detectAndCompute(", "akaze::Akazer::fastDetectAndCompute(",
                "akaze::hLowPass(", "akaze::hDownWithSmooth(", "akaze::hScharrContrast(", "akaze::hFlow(", "akaze::hNldStep(",
                "akaze::hHessianDeterminant(", "akaze::hMatch(", "fastakaze::hConv2dR2(", "fastakaze::hLowPass(", "fastakaze::hDownWithSmooth(",
                "fastakaze::hScharrContrast(", "fastakaze::hFlow(", "fastakaze::hNldStep(", "fastakaze::hHessianDeterminant(",
                "fed_tau_by_process_time(", "fed_tau_by_cycle_time(", "fed_tau_internal(", "fed_is_prime_internal(",
                "setMaxNumPoints(", "getPointCounter(", "setCompareIndices("):
        assert sym in out, sym
    if not B.have_dropin():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    D = B.dropin()
    assert D.dropin_sizeof_point() == 104
    buf = (C.c_float * 16)()
    assert D.dropin_fed_tau(0.53, 1, 0.25, 1, buf, 16) == 3           # host-only entry point of the surface (fed.h)
    assert os.access(B.DROPIN_EXE, os.X_OK)


def test_dropin_headers_compile_against_reference_usage():
    """A translation unit that uses the surface exactly as the reference's main.cpp:170-232 does must compile."""
    src = os.path.join(ROOT, "tests", "cpp", "dropin_usage.cpp")
    inc = ["-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include"]
    subprocess.check_call(["g++", "-std=c++14", "-fsyntax-only"] + inc + [src])


# ---- golden constants --------------------------------------------------------------------------
STEPS5 = [3, 3, 4, 4, 5, 6, 7, 8, 10, 12, 14, 17, 20, 24, 29, 34, 40, 48, 57]


def _level_times(noct, S=4, soffset=np.float32(1.6)):
    last = np.float32(0.5 * float(soffset) * float(soffset))
    out = []
    for o in range(noct):
        for j in range(S):
            if o == 0 and j == 0:
                continue
            es = np.float32(soffset * np.float32(np.power(np.float32(2), np.float32(np.float32(j) / np.float32(S) + np.float32(o)))))
            cur = np.float32(np.float32(0.5) * es * es)
            out.append(np.float32(cur - last))
            last = cur
    return out


def test_fed_schedule_matches_golden_and_reference():
    ts = _level_times(5)
    L = B.oracle()
    all_tau = []
    for T, n_expect in zip(ts, STEPS5):
        mine = ab().fed_tau(float(T), 1, 0.25, True)
        buf = (C.c_float * 128)()
        n = L.orc_fed_tau(float(T), 1, 0.25, 1, buf, 128)
        orc = np.array(buf[:n], dtype=np.float32)
        assert len(mine) == n == n_expect
        assert np.array_equal(mine.view(np.uint32), orc.view(np.uint32))
        if B.have_ref():
            rb = (C.c_float * 128)()
            rn = B.ref().ref_fed_tau(float(T), 1, 0.25, 1, rb, 128)
            assert rn == n and np.array_equal(np.array(rb[:rn], dtype=np.float32).view(np.uint32), mine.view(np.uint32))
        all_tau.append(mine)
        assert abs(float(mine.sum()) - float(T)) <= 1e-4 * max(1.0, float(T))       # a FED cycle integrates to its stopping time
    assert sum(STEPS5[:15]) == 166 and sum(STEPS5) == 345
    np.testing.assert_allclose(all_tau[0], [0.06973, 0.10842, 0.35204], atol=1e-5)
    np.testing.assert_allclose(all_tau[3], [0.14996, 0.96147, 0.11597, 0.27221], atol=1e-5)
    # FNV-1a-64 over the bytes of all tau, levels (0,1), (0,2), ... in order, concatenated as little-endian f32 -- the values
    # of the compiled reference fed.cpp (oracle/_ref) in this container, which the loop above bit-compares directly.  SURVEY
    # App. C lists other hashes (9022012315eb087a / 7eff3dcd05c48ec3) from a survey-time probe whose byte stream was not
    # recorded: neither f32 nor f64 evolution times, reordering on or off, reproduces them, while its step counts and first
    # cycles (asserted above) do agree.  The direct comparison with fed.cpp is the authoritative check.
    assert fnv1a64(np.concatenate(all_tau[:15]).astype("<f4").tobytes()) == 0xd960f8ec79e72c1c
    assert fnv1a64(np.concatenate(all_tau).astype("<f4").tobytes()) == 0x3e38774bab652df9
    # unordered variant and reference cross-check on arbitrary times
    rng = np.random.default_rng(0)
    for T in rng.uniform(0.05, 300, size=40):
        for reorder in (0, 1):
            mine = ab().fed_tau(float(T), 1, 0.25, bool(reorder))
            buf = (C.c_float * 256)()
            n = L.orc_fed_tau(float(T), 1, 0.25, reorder, buf, 256)
            assert np.array_equal(mine.view(np.uint32), np.array(buf[:n], dtype=np.float32).view(np.uint32))
            if B.have_ref():
                rb = (C.c_float * 256)()
                rn = B.ref().ref_fed_tau(float(T), 1, 0.25, reorder, rb, 256)
                assert np.array_equal(np.array(rb[:rn], dtype=np.float32).view(np.uint32), mine.view(np.uint32))


def test_gaussian_taps_golden():
    t2 = ab().gauss_taps(1.0, 2).view(np.uint32)
    assert [hex(v) for v in t2] == ["0x3ece2433", "0x3e7a0fea", "0x3d5f2f86"]
    var0 = float(np.float32(1.6) * np.float32(1.6))          # soffset*soffset in float (akaze.cpp:327), not 2.56
    t4 = ab().gauss_taps(var0, 4).view(np.uint32)
    assert [hex(v) for v in t4] == ["0x3e803503", "0x3e52eba8", "0x3deaca38", "0x3d30d86c", "0x3c3441c3"]
    buf = (C.c_float * 5)()
    B.oracle().orc_gauss_taps(var0, 4, buf)
    assert np.array_equal(np.array(buf[:], dtype=np.float32).view(np.uint32), t4)


def test_comparison_table_golden():
    c1, c2 = ab().compare_indices()
    assert list(zip(c1[:7], c2[:7])) == [(0, 3), (0, 6), (0, 9), (3, 6), (3, 9), (6, 9), (1, 4)]
    assert (c1[126], c2[126]) == (39, 42)
    assert list(zip(c1[-3:], c2[-3:])) == [(80, 83), (80, 86), (83, 86)]
    # independent restatement of akazed.cu:65-159: nine blocks (grid x channel), pairs (j, i > j) in row-major order
    e1, e2 = [], []
    for lo, hi in ((0, 4), (4, 13), (13, 29)):
        for ch in range(3):
            for j in range(lo, hi - 1):
                for i in range(j + 1, hi):
                    e1.append(3 * j + ch)
                    e2.append(3 * i + ch)
    assert len(e1) == 486 == 3 * (6 + 36 + 120)
    assert np.array_equal(c1, e1) and np.array_equal(c2, e2)
    assert c2.max() == 86                                        # 29 cells x 3 channels = values 0..86
    a, b = (C.c_int * 488)(), (C.c_int * 488)()
    B.oracle().orc_compare_indices(a, b)
    assert np.array_equal(np.array(a[:486]), c1) and np.array_equal(np.array(b[:486]), c2)


def test_schedule_of_the_1080p_configuration():
    P = B.OraclePyramid(1920, 1080)
    assert P.levels == 16
    dims = [P.dims(l) for l in range(16)]
    assert [d["nsteps"] for d in dims] == [0] + STEPS5[:15]
    assert [d["sigma_size"] for d in dims] == [2, 3, 3, 4] * 4
    assert [(d["w"], d["h"]) for d in dims[::4]] == [(1920, 1080), (960, 540), (480, 270), (240, 135)]
    np.testing.assert_allclose([d["size"] for d in dims[:4]], [2.4, 2.8541, 3.3941, 4.0363], atol=1e-4)
    assert sum(d["w"] * d["h"] for d in dims) == 11016000            # level-pixels per frame
    assert sum(d["w"] * d["h"] * d["nsteps"] for d in dims) == 40759200
    P.close()
    P = B.OraclePyramid(640, 480)          # 640x480, 4 octaves requested: octave 3 would be 80x60 -> dropped (h < 80)
    assert P.levels == 12
    P.close()


def test_opencv_export_host_only():
    """akz_keypoints_to_opencv / akz_matches_to_opencv are host-only format conversions (no device needed)."""
    k = np.zeros(3, dtype=ab().KEYPOINT_DTYPE)
    k["x"], k["y"], k["size"], k["angle"], k["layer"], k["response"] = [1.5, 2.5, 3.5], [4, 5, 6], [2.4, 2.854, 4.036], [0, np.pi / 2, np.pi], [0, 5, 14], [.1, .2, .3]
    o = ab().keypoints_to_opencv(k, 4)
    assert np.allclose(o[:, 0], [1.5, 2.5, 3.5]) and np.allclose(o[:, 2], [2.4, 2.854 * 2, 4.036 * 8]) and np.allclose(o[:, 3], [0, 90, 180])
    assert o[:, 5].tolist() == [0, 1, 3] and o[:, 6].tolist() == [0, 1, 2]
    m = np.array([[5, 10, -1, 0], [-1, -1, 0, 0], [7, 95, 3, 100]], dtype=np.int32)
    assert ab().matches_to_opencv(m).tolist() == [[0, 5, 10], [2, 7, 95]]


def test_reference_point_layout():
    assert B.REF_POINT.itemsize == 104
    assert [B.REF_POINT.fields[n][1] for n in ("x", "y", "octave", "response", "size", "angle", "features", "match", "distance", "match_x", "match_y")] == \
        [0, 4, 8, 12, 16, 20, 24, 88, 92, 96, 100]
    if B.have_ref():
        assert B.ref().ref_sizeof_point() == 104
    assert ab().KEYPOINT_DTYPE.itemsize == 32 and ab().MATCH_DTYPE.itemsize == 16


# ---- oracle behaviour ----------------------------------------------------------------------------------------------
def test_oracle_stage_properties():
    L = B.oracle()
    w, h = 97, 61
    rng = np.random.default_rng(1)
    img = rng.random((h, w), dtype=np.float32)
    out = np.zeros_like(img)
    # symmetric kernel + commutative pair sums: blurring the mirrored image gives the mirrored blur, bit for bit
    L.orc_lowpass(B._p(img), B._p(out), w, h, w, 2.56, 9)
    flip = np.ascontiguousarray(img[::-1, ::-1])
    out2 = np.zeros_like(img)
    L.orc_lowpass(B._p(flip), B._p(out2), w, h, w, 2.56, 9)
    assert np.array_equal(out.view(np.uint32), out2[::-1, ::-1].view(np.uint32))
    assert abs(out.mean() - img.mean()) < 2e-3
    # zero conductance freezes the diffusion; constant image is a fixed point
    g0 = np.zeros_like(img)
    nxt = np.zeros_like(img)
    L.orc_nld_step(B._p(img), B._p(g0), B._p(nxt), 0.25, w, h, w)
    assert np.array_equal(img, nxt)
    const = np.full_like(img, 0.37)
    g1 = np.ones_like(img)
    L.orc_nld_step(B._p(const), B._p(g1), B._p(nxt), 5.0, w, h, w)
    assert np.array_equal(const, nxt)
    # stable step (0.5*tau*sum(g0+gn) = 0.8 < 1) is a convex combination: discrete maximum principle
    L.orc_nld_step(B._p(img), B._p(g1), B._p(nxt), 0.2, w, h, w)
    assert nxt.max() <= img.max() + 1e-6 and nxt.min() >= img.min() - 1e-6 and nxt.std() < img.std()
    # PM_G2 conductance lies in (0, 1]
    fl = np.zeros_like(img)
    L.orc_flow(B._p(img), B._p(fl), 1, 0.05, w, h, w)
    assert fl.max() <= 1.0 and fl.min() > 0.0
    # derivative filters annihilate constants and reproduce a ramp's slope
    ramp = np.tile(np.arange(w, dtype=np.float32) * 0.5, (h, 1))
    lx, ly, det = np.zeros_like(ramp), np.zeros_like(ramp), np.zeros_like(ramp)
    L.orc_hessian(B._p(ramp), B._p(lx), B._p(ly), B._p(det), 2, w, h, w)
    np.testing.assert_allclose(lx[10:-10, 10:-10], 0.5 * 2 * 2 * (0.09375 * 2 + 0.3125) , rtol=1e-5)   # 2*step*slope*(2*fac1+fac2)
    assert np.abs(ly[10:-10, 10:-10]).max() < 1e-6


def test_oracle_pipeline_properties():
    w, h = 320, 240
    img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=4))
    kps = B.oracle_detect_and_compute(img, noctaves=2)
    assert 50 < len(kps) < 5000
    # raster order, borders respected, 486-bit descriptors with zero padding
    key = kps["iy"].astype(np.int64) * w + kps["ix"]
    assert np.all(np.diff(key) > 0)
    assert kps["ix"].min() >= 28 and kps["iy"].min() >= 28 and kps["ix"].max() < w - 28 and kps["iy"].max() < h - 28
    assert not kps["desc"][:, 61:].any() and not (kps["desc"][:, 60] & 0xC0).any()
    assert np.all((kps["angle"] >= 0) & (kps["angle"] < 2 * np.pi + 1e-6))
    assert np.abs(kps["x"] - kps["ix"]).max() <= 2.0 and np.abs(kps["y"] - kps["iy"]).max() <= 2.0
    # radius NMS: no stronger keypoint inside a keypoint's own disc
    for i in range(0, len(kps), 7):
        d2 = (kps["ix"] - kps["ix"][i]) ** 2 + (kps["iy"] - kps["iy"][i]) ** 2
        inside = (d2 < int(kps["size"][i] ** 2)) & (d2 > 0)
        assert not np.any(kps["response"][inside] > kps["response"][i])
    # run-to-run identical, threads do not change bits
    k1 = B.oracle_detect_and_compute(img, noctaves=2, threads=1)
    assert k1.tobytes() == kps.tobytes()
    # an injected contrast factor is honoured
    k2 = B.oracle_detect_and_compute(img, noctaves=2, kcontrast_override=0.5)
    assert len(k2) != len(kps) or k2.tobytes() != kps.tobytes()


def test_oracle_matchers():
    q = B.random_descriptors(40, 1)
    t = B.random_descriptors(300, 2)
    t[17] = q[0]                                  # unique best -> accepted
    t[40] = q[1]; t[56] = q[1]                    # tie inside stride 8 -> accepted, first index
    t[70] = q[2]; t[75] = q[2]                    # tie across strides -> rejected
    far = q[3].copy(); far[:20] ^= 0xFF           # distance 160 -> rejected by the < 96 gate
    t[90] = far
    r = B.oracle_match(q, t, "compat")
    assert tuple(r[0]) == (17, 0) and tuple(r[1]) == (40, 0) and tuple(r[2]) == (-1, -1)
    assert r[3][0] == -1 or r[3][1] < 96
    k = B.oracle_match(q, t, "knn2")
    assert tuple(k[1]) == (40, 0, 56, 0) and tuple(k[2]) == (70, 0, 75, 0) and k[0][0] == 17
    cv2 = pytest.importorskip("cv2")
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q[:, :61].copy(), t[:, :61].copy(), k=2)
    for i, (a, b) in enumerate(knn):
        assert int(a.distance) == k[i][1] and int(b.distance) == k[i][3]
        if k[i][1] != k[i][3]:
            assert a.trainIdx == k[i][0]


# ---- golden fixtures generated from the compiled reference on the B200 box -------------------------------------------
GOLD = os.path.join(ROOT, "tests", "golden", "ref_320x240.npz")


@pytest.mark.skipif(not os.path.exists(GOLD), reason="reference golden fixture not generated yet (tests/golden/make_ref_golden.py)")
def test_oracle_against_reference_golden_fixture():
    g = np.load(GOLD)
    img = B.u8_to_unit(B.synth_shapes_u8(320, 240, seed=int(g["seed"])))
    assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["img_sha256"])
    P = B.OraclePyramid(320, 240, noctaves=2, kcontrast_override=float(g["kcontrast"]))
    P.build(img)
    assert P.levels == int(g["nlevels"])
    for l in range(P.levels):
        for which, nm in enumerate(("lt", "det", "lx", "ly")):
            mine = P.plane(l, which)
            assert hashlib.sha256(mine.tobytes()).hexdigest() == str(g[f"sha_{l}_{nm}"]), f"level {l} {nm}"
            assert np.array_equal(mine[::8, ::8].view(np.uint32), g[f"sub_{l}_{nm}"].view(np.uint32))
    kp = P.detect()
    ref_xy = np.stack([g["kp_x"], g["kp_y"]], 1)
    from scipy.spatial import cKDTree
    d, j = cKDTree(np.stack([kp["x"], kp["y"]], 1)).query(ref_xy)
    assert (d <= 1e-4).mean() >= 0.99 and abs(len(kp) - len(ref_xy)) <= max(2, 0.01 * len(ref_xy))
    kp = P.describe(kp)
    same = d <= 1e-4
    da = np.abs(kp["angle"][j][same] - g["kp_angle"][same])
    da = np.minimum(da, 2 * np.pi - da)
    assert (da <= 1e-3).mean() >= 0.99
    # descriptors: libm cos/sin vs MUFU and the last bits of the angle can move a sample by one pixel; demand near-identity
    bits = np.unpackbits(kp["desc"][j][same][:, :61] ^ g["kp_desc"][same], axis=1).sum(axis=1)
    assert (bits == 0).mean() >= 0.90 and bits.mean() <= 1.0
    P.close()


GOLD_FAST = os.path.join(ROOT, "tests", "golden", "ref_fast_320x240.npz")


@pytest.mark.skipif(not os.path.exists(GOLD_FAST), reason="integer-pipeline golden fixture not generated yet (tests/golden/make_ref_golden.py)")
def test_fast_oracle_against_reference_golden_fixture():
    """oracle/fast_oracle.py (numpy restatement of Akazer::fastDetect) against the compiled reference: every int plane by
    hash, the integer contrast-factor rule, and the keypoint set with its refined positions."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fast_oracle as FO
    g = np.load(GOLD_FAST)
    img8 = B.synth_shapes_u8(320, 240, seed=int(g["seed"]))
    assert hashlib.sha256(img8.tobytes()).hexdigest() == str(g["img_sha256"])
    lv, k0 = FO.build(img8, noctaves=2, k_override=int(g["kcontrast"]))
    assert len(lv) == int(g["nlevels"])
    for l, L in enumerate(lv):
        for nm, key in (("lt", "Lt"), ("det", "det"), ("lx", "Lx"), ("ly", "Ly")):
            a = np.ascontiguousarray(L[key]).astype(np.int32)
            assert np.array_equal(a[::8, ::8], g[f"sub_{l}_{nm}"]), f"level {l} {nm}"
            assert hashlib.sha256(a.tobytes()).hexdigest() == str(g[f"sha_{l}_{nm}"]), f"level {l} {nm}"
    kp = FO.detect(lv)
    ref = sorted(zip(g["kp_layer"].tolist(), g["kp_y"].view(np.uint32).tolist(), g["kp_x"].view(np.uint32).tolist()))
    mine = sorted(zip(kp["layer"].tolist(), kp["y"].view(np.uint32).tolist(), kp["x"].view(np.uint32).tolist()))
    assert mine == ref and len(ref) > 50
    # the reference's own maximum reduction is racy and partial (App. B-1): its k is injected above; ours uses the true maximum
    assert FO.contrast(FO.conv(img8, 1.0, 2)) > 0


# ---- multi-rank host logic (gloo, world_size 2) -------------------------------------------------------------------------
def _np_partial(q, t, base, mode):
    x = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(axis=2).astype(np.int32)    # (nq, nt)
    out = np.zeros((len(q), 4), dtype=np.int32)
    idx = np.arange(len(t)) + base
    for i in range(len(q)):
        d = x[i]
        if len(t) == 0:
            out[i] = (-1, -1, -1 if mode == 1 else 0, -1 if mode == 1 else 0)
            continue
        if mode == 1:
            order = np.lexsort((idx, d))
            out[i, 0], out[i, 1] = idx[order[0]], d[order[0]]
            if len(t) > 1:
                out[i, 2], out[i, 3] = idx[order[1]], d[order[1]]
            else:
                out[i, 2], out[i, 3] = -1, -1
        else:
            m = d.min()
            hit = idx[d == m]
            out[i] = (hit.min(), m, int(np.bitwise_or.reduce(1 << (hit & 15))), 0)
    return out


def _np_merge(parts, mode):
    world, nq, _ = parts.shape
    out = np.zeros((nq, 4), dtype=np.int32)
    for i in range(nq):
        if mode == 1:
            c = [(parts[p, i, 1], parts[p, i, 0]) for p in range(world) if parts[p, i, 0] >= 0] + \
                [(parts[p, i, 3], parts[p, i, 2]) for p in range(world) if parts[p, i, 2] >= 0]
            c.sort()
            out[i] = (c[0][1], c[0][0], c[1][1], c[1][0]) if len(c) > 1 else (c[0][1], c[0][0], -1, -1)
        else:
            ok = [p for p in range(world) if parts[p, i, 0] >= 0]
            m = min(parts[p, i, 1] for p in ok)
            best = [p for p in ok if parts[p, i, 1] == m]
            mask = 0
            for p in best:
                mask |= int(parts[p, i, 2])
            imin = min(parts[p, i, 0] for p in best)
            acc = bin(mask).count("1") == 1 and m < 96
            out[i] = (imin if acc else -1, m if acc else -1, mask, 0)
    return out


def _gloo_worker(rank, world, port, q, t, mode, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
    from akaze_b200.distributed import shard_bounds, match_sharded
    lo, hi = shard_bounds(len(t), world, rank)
    res = match_sharded(torch.from_numpy(q), torch.from_numpy(t[lo:hi]), lo, mode,
                        lambda q_, t_, b_, m_: torch.from_numpy(_np_partial(q_.numpy(), t_.numpy(), b_, m_)),
                        lambda parts, m_: torch.from_numpy(_np_merge(parts.numpy(), m_)))
    ret[rank] = res.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_train_sharded_matching_world_size_2_gloo(mode):
    import torch.multiprocessing as mp
    q = B.random_descriptors(60, 3)
    t = B.random_descriptors(333, 4)
    t[5] = q[0]; t[200] = q[0]                 # tie across shards and strides
    t[100] = q[1]; t[292] = q[1]               # tie across shards inside stride 4
    t[310] = q[2]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + mode
    mp.spawn(_gloo_worker, args=(2, port, q, t, mode, ret), nprocs=2, join=True)
    exp = B.oracle_match(q, t, "compat" if mode == 0 else "knn2")
    for r in range(2):
        got = ret[r]
        assert np.array_equal(got[:, :exp.shape[1]], exp), (mode, r)


def test_frame_and_train_sharding_bounds():
    from akaze_b200.distributed import shard_bounds
    for n in (0, 1, 7, 256, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_chunk_plan_of_the_host_pipeline():
    """akz_plan_chunks: the chunks of a batch are contiguous, cover it exactly and never exceed max_batch; the host pipeline
    ramps up (B/8, B/4, B/2, B, ...) so that its first upload is short, the device path uses full chunks."""
    L = ab().lib()
    cap = 4096
    starts, sizes = (C.c_int * cap)(), (C.c_int * cap)()
    for n in (0, 1, 3, 5, 11, 64, 100, 256, 1000):
        for B_ in (1, 2, 8, 32):
            for ramp in (0, 1):
                k = L.akz_plan_chunks(n, B_, ramp, starts, sizes, cap)
                assert k >= 0
                st, sz = list(starts[:k]), list(sizes[:k])
                assert sum(sz) == n and all(0 < s <= B_ for s in sz)
                assert st == [sum(sz[:i]) for i in range(k)]
                if not ramp:
                    assert all(s == B_ for s in sz[:-1])
    k = L.akz_plan_chunks(256, 32, 1, starts, sizes, cap)
    assert list(sizes[:k]) == [4, 8, 16] + [32] * 7 + [4]
    assert L.akz_plan_chunks(-1, 8, 0, starts, sizes, cap) < 0 and L.akz_plan_chunks(8, 0, 0, starts, sizes, cap) < 0


def test_work_decomposition_of_the_tcgen05_matcher():
    """akz_plan_match (host arithmetic of match_tc5.cu): one CTA per SM at most, equal tile shares that cover the linear
    tile space, item ranges that are multiples of 1024 descriptors (8 tiles: the 64 columns of a half tile share their index
    class) and at most 2^19 (13-bit half-tile ordinal), enough slots for every CTA that touches an item."""
    L = ab().lib()
    out = (C.c_int * 5)()
    for nq, nt in ((1, 1), (5, 77), (256, 1024), (257, 1025), (10000, 10000), (10000, 100000), (10000, 1000000),
                   (300, 600000), (1 << 20, 4096), (128, 1 << 22)):
        assert L.akz_plan_match(nq, nt, out) == 0, (nq, nt)
        nparts, nsplit, T, S, grid = list(out)
        nqb = (nq + 255) // 256
        total = nqb * nsplit * T
        assert T % 8 == 0 and T >= 8 and T * 128 <= 1 << 19
        assert nsplit * T * 128 >= nt and (nsplit - 1) * T * 128 < max(nt, 1) + T * 128
        assert 1 <= grid <= 148 and grid * S >= total and (grid - 1) * S < total
        maxslots = nparts // nsplit
        assert nparts == nsplit * maxslots
        for item in range(0, nqb * nsplit, max(1, nqb * nsplit // 50)):
            first, last = (item * T) // S, ((item + 1) * T - 1) // S
            assert last - first + 1 <= maxslots
    assert L.akz_plan_match(0, 10, out) < 0
