"""TEST INFRASTRUCTURE — never imported by the product path.

numpy restatement of the reference's INTEGER ("fast") pipeline up to the refined keypoints: Akazer::fastDetect
(akaze.cpp:506-743) and the fastakaze kernels / wrappers of akazed.cu:2781-4366.  Integer arithmetic is exact on a CPU;
the few float expressions are evaluated in float32 the way the sm_100a build of the reference does (the only contraction,
x * 65536 + 0.5f -> fma, is exact either way because x * 65536 is a power-of-two scaling).  Orientation and descriptors use
MUFU approximations (__expf, __cosf, __sinf) and are checked on the GPU against the compiled reference only.

Pinned by tests/golden/ref_fast_320x240.npz (planes, keypoints of the compiled reference; tests/golden/make_ref_golden.py).
"""
import ctypes as C

import numpy as np

f32 = np.float32
NBINS = 300


def gauss_taps(var, radius):
    """createGaussKernel, akazed.cu:3863-3899: float taps, normalised, then (int)(k * 65536 + 0.5f)."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    denom = f32(1.0) / (f32(2.0) * f32(var))
    k = np.array([libm.expf(f32(-i * i) * denom) for i in range(radius + 1)], dtype=f32)
    ksum = f32(0)
    for i in range(radius + 1):
        ksum = f32(ksum + (k[i] if i == 0 else f32(k[i] + k[i])))
    ksum = f32(1) / ksum
    k = (k * ksum).astype(f32)
    return (k * f32(65536) + f32(0.5)).astype(np.int32)


def _pad(a, r):
    return np.pad(a, r, mode="reflect")


def conv(src, var, radius):
    """gConv2d<R> / gConv2dR2 (akazed.cu:2786-3076): rows then columns, each (sum) >> 16; reflect-101 borders."""
    k = gauss_taps(var, radius).astype(np.int64)
    a = src.astype(np.int64)
    h, w = a.shape
    p = np.pad(a, ((0, 0), (radius, radius)), mode="reflect")
    row = k[0] * p[:, radius:radius + w]
    for i in range(1, radius + 1):
        row = row + k[i] * (p[:, radius - i:radius - i + w] + p[:, radius + i:radius + i + w])
    row = (row.astype(np.int32) >> 16).astype(np.int64)
    p = np.pad(row, ((radius, radius), (0, 0)), mode="reflect")
    col = k[0] * p[radius:radius + h]
    for i in range(1, radius + 1):
        col = col + k[i] * (p[radius - i:radius - i + h] + p[radius + i:radius + i + h])
    return (col.astype(np.int32) >> 16)


def radius_from_ksz(ksz):
    return 2 if ksz <= 5 else 3 if ksz <= 7 else 4 if ksz <= 9 else 5


def down_with_smooth(src):
    """gDownWithSmooth (akazed.cu:3143-3205): dst = src(2x, 2y); smooth on the coarse lattice, reflect in SOURCE coordinates."""
    sh, sw = src.shape
    dh, dw = sh >> 1, sw >> 1
    k = gauss_taps(1.0, 2).astype(np.int64)

    def sref(c, m):
        c = np.abs(c)
        return np.where(c < m, c, 2 * m - 2 - c)
    xs = [sref(2 * np.arange(dw) + o, sw) for o in (-4, -2, 0, 2, 4)]
    ys = [sref(2 * np.arange(dh) + o, sh) for o in (-4, -2, 0, 2, 4)]
    a = src.astype(np.int64)
    rows = []
    for r in range(5):
        R = a[ys[r]]
        v = k[0] * R[:, xs[2]] + k[1] * (R[:, xs[1]] + R[:, xs[3]]) + k[2] * (R[:, xs[0]] + R[:, xs[4]])
        rows.append((v.astype(np.int32) >> 16).astype(np.int64))
    sm = k[0] * rows[2] + k[1] * (rows[1] + rows[3]) + k[2] * (rows[0] + rows[4])
    return src[::2, ::2][:dh, :dw].astype(np.int32).copy(), (sm.astype(np.int32) >> 16)


def _nb(a, step):
    h, w = a.shape
    p = _pad(a.astype(np.int32), step)
    s = step
    g = lambda dy, dx: p[s + dy:s + dy + h, s + dx:s + dx + w]
    return dict(ul=g(-s, -s), uc=g(-s, 0), ur=g(-s, s), cl=g(0, -s), cr=g(0, s), ll=g(s, -s), lc=g(s, 0), lr=g(s, s))


def scharr(a):
    n = _nb(a, 1)
    dx = 10 * (n["cr"] - n["cl"]) + 3 * (n["ur"] + n["lr"] - n["ul"] - n["ll"])
    dy = 10 * (n["lc"] - n["uc"]) + 3 * (n["ll"] + n["lr"] - n["ul"] - n["ur"])
    return dx.astype(np.int32), dy.astype(np.int32)


def contrast(smooth, per=0.7):
    """hScharrContrast (akazed.cu:4094-4169) with the TRUE maximum (App. B-1), in-image histogram (B-3)."""
    dx, dy = scharr(smooth)
    mag = (np.sqrt((dx * dx + dy * dy).astype(f32)) + f32(0.5)).astype(np.int32)          # akazed.cu:3230
    hmax = max(1, int(mag.max()))
    factor = int(f32(f32(f32(NBINS) / f32(hmax)) * f32(65536)) + f32(0.5))
    hi = np.minimum((mag * np.int32(factor)) >> 16, NBINS - 1)
    hist = np.bincount(hi.ravel(), minlength=NBINS)
    thresh = int(f32(smooth.size - hist[0]) * f32(per))
    cum, k = 0, 1
    while k < NBINS:
        if cum >= thresh:
            break
        cum += int(hist[k]); k += 1
    return k * hmax // NBINS


def flow(smooth, k, type=1):
    """gFlowNaive (akazed.cu:3406-3446), PM_G2 / CHARBONNIER (the exp-based ones need MUFU)."""
    dx, dy = scharr(smooth)
    ikc = f32(1.0) / f32(k * k)
    dif2 = (dx * dx + dy * dy).astype(np.uint32).astype(f32) * ikc
    if type == 1:
        g = f32(1.0) / (f32(1.0) + dif2)
    elif type == 3:
        g = f32(1.0) / np.sqrt(f32(1.0) + dif2)
    else:
        raise NotImplementedError("exp-based conductances use MUFU.EX2 on the device")
    return (g.astype(f32) * f32(65536) + f32(0.5)).astype(np.int32)


def nld_step(L, g, tau):
    """gNldStepNaive (akazed.cu:3449-3473); int32 wrap-around like the device."""
    stepfac = np.int32(int(f32(f32(f32(0.5) * f32(tau)) * f32(65536)) + f32(0.5)))
    Lp, gp = _pad(L.astype(np.int32), 1), _pad(g.astype(np.int32), 1)
    h, w = L.shape
    c = lambda p, dy, dx: p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    L0, g0 = c(Lp, 0, 0), c(gp, 0, 0)
    with np.errstate(over="ignore"):
        s = (g0 + c(gp, 0, 1)) * (c(Lp, 0, 1) - L0) + (g0 + c(gp, 0, -1)) * (c(Lp, 0, -1) - L0) + \
            (g0 + c(gp, 1, 0)) * (c(Lp, 1, 0) - L0) + (g0 + c(gp, -1, 0)) * (c(Lp, -1, 0) - L0)
        step = s.astype(np.int32) >> 16
        return (((stepfac * step).astype(np.int32)) >> 16) + L0


def hessian(smooth, step):
    """gDerivate + gHessianDeterminant (akazed.cu:3339-3403), ifac = (int)(fac * 65536 + 0.5f) = 6144, 20480."""
    w = f32(10.0) / f32(3.0)
    fac1 = f32(1.0) / (f32(2.0) * (w + f32(2.0)))
    fac2 = w * fac1
    i1, i2 = np.int32(int(fac1 * f32(65536) + f32(0.5))), np.int32(int(fac2 * f32(65536) + f32(0.5)))
    n = _nb(smooth, step)
    with np.errstate(over="ignore"):
        lx = (i1 * (n["ur"] + n["lr"] - n["ul"] - n["ll"]) + i2 * (n["cr"] - n["cl"])) >> 16
        ly = (i1 * (n["lr"] + n["ll"] - n["ur"] - n["ul"]) + i2 * (n["lc"] - n["uc"])) >> 16
        a, b = _nb(lx, step), _nb(ly, step)
        dxx = (i1 * (a["ur"] + a["lr"] - a["ul"] - a["ll"]) + i2 * (a["cr"] - a["cl"])) >> 16
        dxy = (i1 * (a["lr"] + a["ll"] - a["ur"] - a["ul"]) + i2 * (a["lc"] - a["uc"])) >> 16
        dyy = (i1 * (b["lr"] + b["ll"] - b["ur"] - b["ul"]) + i2 * (b["lc"] - b["uc"])) >> 16
        det = dxx * dyy - dxy * dxy
    return lx.astype(np.int32), ly.astype(np.int32), det.astype(np.int32)


def schedule(w, h, noctaves=4, S=4, soffset=1.6, derivative_factor=1.5):
    """Level list of akaze.cpp:204-237 / :268-363 with FED taus from the oracle C library (fed.cpp)."""
    import bindings as B
    libm = C.CDLL("libm.so.6")
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    L = B.oracle()
    levels, last = [], f32(0.5) * f32(soffset) * f32(soffset)
    ow, oh = w, h
    for i in range(noctaves):
        if i and ((ow >> 1) < 80 or (oh >> 1) < 80):
            break
        if i:
            ow, oh = ow >> 1, oh >> 1
        for j in range(S):
            if i == 0 and j == 0:
                size, tau = f32(soffset) * f32(derivative_factor), []
            else:
                es = f32(f32(soffset) * f32(libm.powf(f32(2), f32(f32(j) / f32(S)) + f32(i))))
                cur = f32(f32(0.5) * es) * es
                buf = (C.c_float * 256)()
                n = L.orc_fed_tau(f32(cur - last), 1, f32(0.25), 1, buf, 256)
                tau = [f32(buf[k]) for k in range(n)]
                size = f32(f32(es * f32(derivative_factor)) / f32(1 << i))
                last = cur
            levels.append(dict(octave=i, sub=j, w=ow, h=oh, size=f32(size), sigma_size=int(f32(size + f32(0.5))), tau=tau))
    return levels


def build(img8, noctaves=4, S=4, per=0.7, soffset=1.6, k_override=0, diffusivity=1):
    """Integer scale space of Akazer::fastDetect.  Returns (levels, k) with planes Lt, det, Lx, Ly per level."""
    h, w = img8.shape
    lv = schedule(w, h, noctaves, S, soffset)
    var0 = f32(soffset) * f32(soffset)
    ksz0 = int(2 * np.ceil((f32(soffset) - f32(0.8)) / f32(0.3)) + 3)
    smooth = conv(img8, 1.0, 2)
    k0 = k_override if k_override > 0 else contrast(smooth, per)
    k = k0
    lt = conv(img8, var0, radius_from_ksz(ksz0))
    lx, ly, det = hessian(lt, lv[0]["sigma_size"])
    lv[0].update(Lt=lt, Lx=lx, Ly=ly, det=det)
    for l in range(1, len(lv)):
        L = lv[l]
        if L["sub"] == 0:
            k = int(f32(f32(k) * f32(0.75)) + f32(0.5))
            cur, smooth = down_with_smooth(lv[l - S]["Lt"])
        else:
            cur = lv[l - 1]["Lt"]
            smooth = conv(cur, 1.0, 2)
        g = flow(smooth, k, diffusivity)
        for t in L["tau"]:
            cur = nld_step(cur, g, t)
        lx, ly, det = hessian(smooth, L["sigma_size"])
        L.update(Lt=cur.astype(np.int32), Lx=lx, Ly=ly, det=det)
    return lv, k0


def detect(lv, S=4, threshold=65):
    """gCalcExtremaMap (serial merge), gNmsRNaive incl. its centre-row indexing, gRefine (akazed.cu:3476-3646).
    Returns a structured array (x, y, layer, size, ix, iy, response) in raster order."""
    W, H = lv[0]["w"], lv[0]["h"]
    smax = f32(10.0 * float(np.sqrt(f32(2.0))))
    resp = np.zeros((H, W), dtype=np.int64)
    layer = np.full((H, W), -1, dtype=np.int32)
    psz = 10000.0
    for l, L in enumerate(lv):
        o, w, h, det = L["octave"], L["w"], L["h"], L["det"]
        border = f32(smax * f32(L["sigma_size"]))
        b0 = int(f32(smax * f32(lv[o * S]["sigma_size"])))
        psz = min(psz, float(f32(smax * f32(lv[o * S]["sigma_size"]))) * (1 << o))
        ys, xs = np.mgrid[b0:h - 1, b0:w - 1]
        ok = ((xs.astype(f32) - border + f32(0.5)).astype(np.int32) - 1 >= 0) & ((xs.astype(f32) + border + f32(0.5)).astype(np.int32) + 1 < w) & \
             ((ys.astype(f32) - border + f32(0.5)).astype(np.int32) - 1 >= 0) & ((ys.astype(f32) + border + f32(0.5)).astype(np.int32) + 1 < h)
        v = det[ys, xs]
        ok &= v > threshold
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dy or dx:
                    ok &= v > det[ys + dy, xs + dx]
        for y, x in zip(ys[ok], xs[ok]):
            oy, ox = y << o, x << o
            if resp[oy, ox] < det[y, x]:
                resp[oy, ox] = det[y, x]; layer[oy, ox] = l
    psz = int(psz)
    out = []
    cy, cx = np.nonzero(layer[psz:H - psz, psz:W - psz] >= 0)
    for iy, ix in zip(cy + psz, cx + psz):
        l = layer[iy, ix]
        fsz = lv[l]["size"]
        isz, sq = int(f32(fsz + f32(0.5))), int(f32(fsz * fsz))
        rc, kill = resp[iy, ix], False
        for i in range(-isz, isz + 1):
            for j in range(-isz, isz + 1):
                if (i == 0 and j == 0) or i * i + j * j >= sq:
                    continue
                jj = j - 1 if (i == 0 and j > 0) else j                        # akazed.cu:3562-3576: `continue` skips new_idx++
                if layer[iy + i, ix + jj] < 0:
                    continue
                rn = resp[iy + i, ix + jj]
                if rn > rc or (rn == rc and i <= 0 and j <= 0):
                    kill = True
                    break
            if kill:
                break
        if kill:
            continue
        L = lv[l]
        o = L["octave"]
        d = L["det"].astype(np.int64)
        y, x = iy >> o, ix >> o
        v2 = d[y, x] + d[y, x]
        gx, gy = (d[y, x + 1] - d[y, x - 1]) >> 1, (d[y + 1, x] - d[y - 1, x]) >> 1
        dxx, dyy = d[y, x + 1] + d[y, x - 1] - v2, d[y + 1, x] + d[y - 1, x] - v2
        dxy = (d[y + 1, x + 1] + d[y - 1, x - 1] - d[y - 1, x + 1] - d[y + 1, x - 1]) >> 2
        i32 = lambda v: int(np.int64(v).astype(np.int32))
        dd = i32(i32(dxx * dyy) - i32(dxy * dxy))
        idd = f32(1.0) / f32(dd) if dd != 0 else f32(0)
        o0 = f32(idd * f32(i32(i32(dxy * gy) - i32(dyy * gx))))
        o1 = f32(idd * f32(i32(i32(dxy * gx) - i32(dxx * gy))))
        if o0 < -1 or o0 > 1 or o1 < -1 or o1 > 1:
            px, py = f32(ix), f32(iy)
        else:
            px, py = f32(f32(1 << o) * f32(f32(x) + o0)), f32(f32(1 << o) * f32(f32(y) + o1))
        out.append((px, py, l, fsz, ix, iy, int(rc)))
    return np.array(out, dtype=[("x", "<f4"), ("y", "<f4"), ("layer", "<i4"), ("size", "<f4"), ("ix", "<i4"), ("iy", "<i4"), ("response", "<i4")])
