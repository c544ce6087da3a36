"""ctypes bindings of the two checkers (TEST INFRASTRUCTURE):
   oracle/liboracle_akaze.so    CPU restatement (oracle/akaze_oracle.c)
   oracle/_ref/libref_akaze.so  the reference itself, compiled for sm_100a with our shim (oracle/ref_shim.cu)
plus the seeded synthetic inputs shared by tests and bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_akaze.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_akaze.so")
REF_DATA = os.path.join(ORACLE_DIR, "_ref", "data")


class OrcOptions(C.Structure):
    _fields_ = [("noctaves", C.c_int), ("max_scale", C.c_int), ("per", C.c_float), ("soffset", C.c_float),
                ("reordering", C.c_int), ("derivative_factor", C.c_float), ("dthreshold", C.c_float),
                ("diffusivity", C.c_int), ("pattern", C.c_int), ("kcontrast_override", C.c_float), ("threads", C.c_int)]


ORC_KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                   ("layer", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("desc", "u1", 64)])

# reference AkazePoint (akaze_structures.h:19-39), 104 bytes
REF_POINT = np.dtype({"names": ["x", "y", "octave", "response", "size", "angle", "features", "match", "distance", "match_x", "match_y"],
                      "formats": ["<f4", "<f4", "<i4", "<f4", "<f4", "<f4", ("u1", 61), "<i4", "<i4", "<f4", "<f4"],
                      "offsets": [0, 4, 8, 12, 16, 20, 24, 88, 92, 96, 100], "itemsize": 104})

_orc = None


def oracle():
    global _orc
    if _orc is not None:
        return _orc
    src = os.path.join(ORACLE_DIR, "akaze_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle_akaze.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(ORACLE_SO)
    f, i, vp = C.c_float, C.c_int, C.c_void_p
    L.orc_default_options.argtypes = [C.POINTER(OrcOptions)]
    L.orc_fed_tau.argtypes = [f, i, f, i, C.POINTER(f), i]
    L.orc_gauss_taps.argtypes = [f, i, C.POINTER(f)]
    L.orc_lowpass.argtypes = [vp, vp, i, i, i, f, i]
    L.orc_down_with_smooth.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.orc_scharr_mag.argtypes = [vp, vp, i, i, i]
    L.orc_contrast_from_mag.argtypes = [vp, i, i, i, f, C.POINTER(f)]
    L.orc_contrast_from_mag.restype = f
    L.orc_flow.argtypes = [vp, vp, i, f, i, i, i]
    L.orc_nld_step.argtypes = [vp, vp, vp, f, i, i, i]
    L.orc_hessian.argtypes = [vp, vp, vp, vp, i, i, i, i]
    L.orc_pyramid_create.argtypes = [i, i, C.POINTER(OrcOptions)]
    L.orc_pyramid_create.restype = vp
    L.orc_pyramid_free.argtypes = [vp]
    L.orc_pyramid_levels.argtypes = [vp]
    L.orc_pyramid_level_dims.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(f), C.POINTER(i)]
    L.orc_pyramid_tau.argtypes = [vp, i, C.POINTER(f), i]
    L.orc_pyramid_plane.argtypes = [vp, i, i]
    L.orc_pyramid_plane.restype = C.POINTER(f)
    L.orc_pyramid_kcontrast.argtypes = [vp]
    L.orc_pyramid_kcontrast.restype = f
    L.orc_pyramid_build.argtypes = [vp, vp, i]
    L.orc_pyramid_detect.argtypes = [vp, vp, i]
    L.orc_pyramid_describe.argtypes = [vp, vp, i, i, vp]
    L.orc_detect_and_compute.argtypes = [vp, i, i, i, C.POINTER(OrcOptions), vp, i, i]
    L.orc_match_compat.argtypes = [vp, i, vp, i, vp]
    L.orc_match_knn2.argtypes = [vp, i, vp, i, vp]
    L.orc_compare_indices.argtypes = [C.POINTER(i), C.POINTER(i)]
    _orc = L
    return L


def orc_options(**kw):
    o = OrcOptions()
    oracle().orc_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _p(a):
    return C.c_void_p(a.ctypes.data)


class OraclePyramid:
    def __init__(self, w, h, **kw):
        self.L = oracle()
        self.opt = orc_options(**kw)
        self.w, self.h = w, h
        self.P = self.L.orc_pyramid_create(w, h, C.byref(self.opt))

    def close(self):
        if self.P:
            self.L.orc_pyramid_free(self.P)
            self.P = None

    def __del__(self):
        self.close()

    @property
    def levels(self):
        return self.L.orc_pyramid_levels(self.P)

    def dims(self, l):
        w, h, p, n, ss = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        sz = C.c_float()
        self.L.orc_pyramid_level_dims(self.P, l, C.byref(w), C.byref(h), C.byref(p), C.byref(n), C.byref(sz), C.byref(ss))
        return dict(w=w.value, h=h.value, pitch=p.value, nsteps=n.value, size=sz.value, sigma_size=ss.value)

    def tau(self, l):
        buf = (C.c_float * 128)()
        n = self.L.orc_pyramid_tau(self.P, l, buf, 128)
        return np.array(buf[:n], dtype=np.float32)

    def build(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self.L.orc_pyramid_build(self.P, _p(img), img.shape[1])

    def plane(self, l, which):
        d = self.dims(l)
        ptr = self.L.orc_pyramid_plane(self.P, l, which)
        a = np.ctypeslib.as_array(ptr, shape=(d["h"], d["pitch"]))
        return a[:, :d["w"]].copy()

    @property
    def kcontrast(self):
        return float(self.L.orc_pyramid_kcontrast(self.P))

    def detect(self, cap=100000):
        kps = np.zeros(cap, dtype=ORC_KP)
        n = self.L.orc_pyramid_detect(self.P, _p(kps), cap)
        return kps[:n].copy()

    def describe(self, kps, with_orientation=True, trig=None):
        kps = np.ascontiguousarray(kps)
        t = np.ascontiguousarray(trig, dtype=np.float32) if trig is not None else None
        self.L.orc_pyramid_describe(self.P, _p(kps), len(kps), int(with_orientation), _p(t) if t is not None else None)
        return kps


def oracle_detect_and_compute(img, describe=True, cap=100000, **kw):
    L = oracle()
    o = orc_options(**kw)
    img = np.ascontiguousarray(img, dtype=np.float32)
    kps = np.zeros(cap, dtype=ORC_KP)
    n = L.orc_detect_and_compute(_p(img), img.shape[1], img.shape[0], img.shape[1], C.byref(o), _p(kps), cap, int(describe))
    return kps[:n].copy()


def oracle_match(q, t, mode="compat"):
    L = oracle()
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    if mode == "compat":
        out = np.zeros((len(q), 2), dtype=np.int32)
        L.orc_match_compat(_p(q), len(q), _p(t), len(t), _p(out))
    else:
        out = np.zeros((len(q), 4), dtype=np.int32)
        L.orc_match_knn2(_p(q), len(q), _p(t), len(t), _p(out))
    return out


# ---- the compiled reference ---------------------------------------------------------------------------
_ref = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(REF_SO)
    f, i, vp = C.c_float, C.c_int, C.c_void_p
    L.ref_set_kcontrast_mode.argtypes = [i, f]
    L.ref_last_kcontrast.restype = f
    L.ref_fed_tau.argtypes = [f, i, f, i, C.POINTER(f), i]
    L.ref_fed_tau_internal.argtypes = [i, f, f, i, C.POINTER(f), i]
    L.ref_setOparam.argtypes = [C.POINTER(i), i]
    L.ref_setMaxNumPoints.argtypes = [i]
    L.ref_readPointCounter.restype = C.c_uint
    L.ref_hLowPass.argtypes = [vp, vp, i, i, i, f, i]
    L.ref_hDownWithSmooth.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.ref_hScharrContrast.argtypes = [vp, vp, f, i, i, i]
    L.ref_hScharrContrast.restype = f
    L.ref_hFlow.argtypes = [vp, vp, i, f, i, i, i]
    L.ref_hNldStep.argtypes = [vp, vp, vp, f, i, i, i]
    L.ref_hHessianDeterminant.argtypes = [vp, vp, vp, i, i, i, i]
    L.ref_hCalcExtremaMap.argtypes = [vp, vp, vp, vp, C.POINTER(f), i, i, f, i, i, i, i]
    L.ref_hNmsR.argtypes = [vp, vp, vp, vp, i, i, i, i, i]
    L.ref_hRefine.argtypes = [vp, i, i, vp, i, i]
    L.ref_hCalcOrient.argtypes = [vp, i, i, vp, i, i]
    L.ref_hDescribe.argtypes = [vp, i, i, vp, i, i, i]
    L.ref_hMatch.argtypes = [vp, i, vp, i]
    L.ref_akazer_create.argtypes = [i, i, i, i, i, f, f, f, i, f, f, i, i]
    L.ref_akazer_create.restype = vp
    L.ref_akazer_destroy.argtypes = [vp]
    L.ref_akazer_detectAndCompute.argtypes = [vp, vp, i, i, i, i, vp, vp, i]
    L.ref_akazer_detect_keep.argtypes = [vp, vp, i, i, i, i, vp, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i)]
    L.ref_cuda_free.argtypes = [vp]
    # integer pipeline
    L.ref_fast_set_kcontrast_mode.argtypes = [i, i]
    L.ref_fast_hConv2dR2_u8.argtypes = [vp, vp, i, i, i, f]
    L.ref_fast_hConv2dR2_i.argtypes = [vp, vp, i, i, i, f]
    L.ref_fast_hLowPass.argtypes = [vp, vp, i, i, i, f, i]
    L.ref_fast_hDownWithSmooth.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.ref_fast_hScharrContrast.argtypes = [vp, vp, f, i, i, i]
    L.ref_fast_hFlow.argtypes = [vp, vp, i, i, i, i, i]
    L.ref_fast_hNldStep.argtypes = [vp, vp, vp, f, i, i, i]
    L.ref_fast_hHessianDeterminant.argtypes = [vp, vp, vp, i, i, i, i]
    L.ref_fast_hCalcExtremaMap.argtypes = [vp, vp, vp, vp, C.POINTER(f), i, i, i, i, i, i, i]
    L.ref_fast_hNmsR.argtypes = [vp, vp, vp, vp, i, i, i, i, i]
    L.ref_fast_hRefine.argtypes = [vp, i, i, vp, i, i]
    L.ref_fast_hCalcOrient.argtypes = [vp, i, i, vp, i, i]
    L.ref_fast_hDescribe.argtypes = [vp, i, i, vp, i, i, i]
    L.ref_akazer_fast_detect_keep.argtypes = [vp, vp, i, i, i, i, vp, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i)]
    L.ref_akazer_fastDetectAndCompute.argtypes = [vp, vp, i, i, i, i, vp, vp, i]
    L.ref_cuMatch.argtypes = [vp, vp, i, vp, i]
    L.ref_scrub_shared_memory.restype = i
    _ref = L
    return L



# ---- the PRODUCT's drop-in C++ surface (include/akaze.h ...) behind tests/cpp/dropin_shim.cu -------------------------
DROPIN_DIR = os.path.join(ROOT, "tests", "cpp", "build")
DROPIN_SO = os.path.join(DROPIN_DIR, "libdropin_shim.so")
DROPIN_EXE = os.path.join(DROPIN_DIR, "dropin_usage")
_dropin = None


def have_dropin():
    return os.path.exists(DROPIN_SO)


def dropin():
    """ctypes handle of the shim that drives akaze::Akazer / initAkazeData / cuMatch / the h* stage functions of the
    product library exactly as the reference's main.cpp does (the twin of ref())."""
    global _dropin
    if _dropin is not None:
        return _dropin
    L = C.CDLL(DROPIN_SO)
    f, i, vp = C.c_float, C.c_int, C.c_void_p
    L.dropin_data_create.argtypes = [i, i, i]
    L.dropin_data_create.restype = vp
    L.dropin_data_free.argtypes = [vp]
    for name in ("dropin_data_num", "dropin_data_max"):
        getattr(L, name).argtypes = [vp]
    for name in ("dropin_data_host", "dropin_data_dev"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.dropin_data_set_num.argtypes = [vp, i]
    L.dropin_akazer_create.argtypes = [i, i, i, i, i, f, f, f, i, f, f, i, i]
    L.dropin_akazer_create.restype = vp
    L.dropin_akazer_destroy.argtypes = [vp]
    L.dropin_akazer_detectAndCompute.argtypes = [vp, vp, vp, i, i, i, i]
    L.dropin_akazer_fastDetectAndCompute.argtypes = [vp, vp, vp, i, i, i, i]
    L.dropin_akazer_time.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.dropin_akazer_time.restype = f
    L.dropin_cuMatch.argtypes = [vp, vp]
    L.dropin_hMatch.argtypes = [vp, vp]
    L.dropin_fed_tau.argtypes = [f, i, f, i, C.POINTER(f), i]
    L.dropin_fed_tau_internal.argtypes = [i, f, f, i, C.POINTER(f), i]
    L.dropin_hLowPass.argtypes = [vp, vp, i, i, i, f, i]
    L.dropin_hDownWithSmooth.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.dropin_hScharrContrast.argtypes = [vp, vp, f, i, i, i]
    L.dropin_hScharrContrast.restype = f
    L.dropin_hFlow.argtypes = [vp, vp, i, f, i, i, i]
    L.dropin_hNldStep.argtypes = [vp, vp, vp, f, i, i, i]
    L.dropin_hHessianDeterminant.argtypes = [vp, vp, vp, i, i, i, i]
    L.dropin_fast_hConv2dR2_u8.argtypes = [vp, vp, i, i, i, f]
    L.dropin_fast_hConv2dR2_i.argtypes = [vp, vp, i, i, i, f]
    L.dropin_fast_hConv2dR2_u8_t.argtypes = [vp, vp, vp, i, i, i, f]
    L.dropin_fast_hConv2dR2_i_t.argtypes = [vp, vp, vp, i, i, i, f]
    L.dropin_fast_hLowPass.argtypes = [vp, vp, i, i, i, f, i]
    L.dropin_fast_hLowPass_t.argtypes = [vp, vp, vp, i, i, i, f, i]
    L.dropin_fast_hDownWithSmooth.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.dropin_fast_hScharrContrast.argtypes = [vp, vp, f, i, i, i]
    L.dropin_fast_hFlow.argtypes = [vp, vp, i, i, i, i, i]
    L.dropin_fast_hNldStep.argtypes = [vp, vp, vp, f, i, i, i]
    L.dropin_fast_hHessianDeterminant.argtypes = [vp, vp, vp, i, i, i, i]
    _dropin = L
    return L


class DropinData:
    """akaze::AkazeData allocated by the product's initAkazeData (akaze.h:12); host / device records as REF_POINT arrays."""

    def __init__(self, max_pts, host=True, dev=True):
        self.L = dropin()
        self.h = self.L.dropin_data_create(max_pts, int(host), int(dev))
        self.max_pts = max_pts

    @property
    def num(self):
        return self.L.dropin_data_num(self.h)

    def host_records(self, n=None):
        n = self.num if n is None else n
        p = self.L.dropin_data_host(self.h)
        if not p or n <= 0:
            return np.zeros(0, dtype=REF_POINT)
        buf = (C.c_uint8 * (n * 104)).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8).copy().view(REF_POINT)

    def dev_records(self, n=None):
        import torch
        n = self.num if n is None else n
        p = self.L.dropin_data_dev(self.h)
        if not p or n <= 0:
            return np.zeros(0, dtype=REF_POINT)
        torch.cuda.synchronize()
        # raw bytes first: a structured-array .copy() copies field by field and leaves the padding bytes (85..87) uninitialised
        return _wrap_device(p, n * 104).cpu().numpy().copy().view(REF_POINT)

    @property
    def dev_ptr(self):
        return self.L.dropin_data_dev(self.h)

    def close(self):
        if self.h:
            self.L.dropin_data_free(self.h)
            self.h = None


class DropinAkazer:
    """akaze::Akazer of the PRODUCT (include/akaze.h) driven like main.cpp:192-205 drives the reference's."""

    def __init__(self, w, h, pitch, noctaves=4, max_scale=4, per=0.7, kcontrast=0.03, soffset=1.6, reordering=True,
                 derivative_factor=1.5, dthreshold=0.001, diffusivity=1, pattern=10):
        self.L = dropin()
        self.w, self.h, self.pitch = w, h, pitch
        self.hnd = self.L.dropin_akazer_create(w, h, pitch, noctaves, max_scale, per, kcontrast, soffset, int(reordering),
                                               derivative_factor, dthreshold, diffusivity, pattern)

    def detect_and_compute(self, img_t, data, desc=True, fast=False):
        import torch
        torch.cuda.synchronize()
        fn = self.L.dropin_akazer_fastDetectAndCompute if fast else self.L.dropin_akazer_detectAndCompute
        return fn(self.hnd, C.c_void_p(img_t.data_ptr()), data.h, self.w, self.h, self.pitch, int(desc))

    def time(self, img_t, data, iters, desc=True, fast=False):
        """ms per call of the synchronous entry point (the reference's own timed loop, main.cpp:199-205)"""
        return float(self.L.dropin_akazer_time(self.hnd, C.c_void_p(img_t.data_ptr()), data.h, self.w, self.h, self.pitch,
                                               int(desc), int(fast), iters))

    def close(self):
        if self.hnd:
            self.L.dropin_akazer_destroy(self.hnd)
            self.hnd = None


def _wrap_device(ptr, nbytes):
    """A uint8 torch view of foreign device memory (no ownership)."""
    import torch

    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda")


def ref_schedule(noct, S=4, soffset=1.6, derivative_factor=1.5):
    """(sizes, sigma_sizes, borders) per octave exactly as akaze.cpp:268-363 computes them (float arithmetic, libm powf)."""
    libm = C.CDLL("libm.so.6")
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    f = np.float32
    smax = f(10.0 * float(np.sqrt(f(2.0))))
    out = []
    for i in range(noct):
        oratio = f(1 << i)
        sizes, ss, borders = [], [], []
        for j in range(S):
            if i == 0 and j == 0:
                size = f(f(soffset) * f(derivative_factor))
            else:
                esigma = f(f(soffset) * f(libm.powf(f(2), f(f(j) / f(S)) + f(i))))
                size = f(f(esigma * f(derivative_factor)) / oratio)
            sg = int(f(size + f(0.5)))
            sizes.append(size); ss.append(sg); borders.append(f(smax * f(sg)))
        out.append((np.array(sizes, dtype=f), ss, np.array(borders, dtype=f)))
    return out


class RefAkazer:
    """The reference's Akazer (akaze.h:19-66) behind the shim; images are torch CUDA tensors (h, pitch)."""

    def __init__(self, w, h, pitch, noctaves=4, max_scale=4, per=0.7, kcontrast=0.03, soffset=1.6, reordering=True,
                 derivative_factor=1.5, dthreshold=0.001, diffusivity=1, pattern=10):
        self.L = ref()
        self.w, self.h, self.pitch, self.S = w, h, pitch, max_scale
        self.hnd = self.L.ref_akazer_create(w, h, pitch, noctaves, max_scale, per, kcontrast, soffset, int(reordering),
                                            derivative_factor, dthreshold, diffusivity, pattern)

    def close(self):
        if self.hnd:
            self.L.ref_akazer_destroy(self.hnd)
            self.hnd = None

    def detect_and_compute(self, img_t, max_pts=100000, desc=True):
        import torch
        pts = torch.zeros(max_pts * 104, dtype=torch.uint8, device=img_t.device)
        torch.cuda.synchronize()
        n = self.L.ref_akazer_detectAndCompute(self.hnd, C.c_void_p(img_t.data_ptr()), self.w, self.h, self.pitch, int(desc),
                                               C.c_void_p(pts.data_ptr()), None, max_pts)
        torch.cuda.synchronize()
        return pts.cpu().numpy().view(REF_POINT)[:n].copy(), pts

    def detect_keep(self, img_t, max_pts=100000, desc=True):
        """Runs the reference pipeline and returns (points, planes) with planes[level][which] as (h, w) arrays,
        which: 0 Lt, 1 det, 2 Lx, 3 Ly (akaze.cpp:315-320)."""
        import torch
        pts = torch.zeros(max_pts * 104, dtype=torch.uint8, device=img_t.device)
        tmem = C.c_void_p()
        oparams = (C.c_int * 64)()
        noct = C.c_int()
        torch.cuda.synchronize()
        n = self.L.ref_akazer_detect_keep(self.hnd, C.c_void_p(img_t.data_ptr()), self.w, self.h, self.pitch, int(desc),
                                          C.c_void_p(pts.data_ptr()), max_pts, C.byref(tmem), oparams, C.byref(noct))
        torch.cuda.synchronize()
        no = noct.value
        osizes = list(oparams[0:no])
        offsets = list(oparams[no:2 * no + 1])
        owhps = [tuple(oparams[2 * no + 1 + 3 * k: 2 * no + 4 + 3 * k]) for k in range(no)]
        total = offsets[no]
        from akaze_b200 import _cudart_memcpy_d2d
        buf = torch.empty(total, dtype=torch.float32, device=img_t.device)
        _cudart_memcpy_d2d(buf.data_ptr(), tmem.value, total * 4)
        host = buf.cpu().numpy()
        self.L.ref_cuda_free(tmem)
        planes = []
        S = self.S
        for o in range(no):
            w, h, p = owhps[o]
            for s in range(S):
                grp = []
                for which in range(4):
                    off = offsets[o] + (which * S + s) * osizes[o]
                    grp.append(host[off:off + osizes[o]].reshape(h, p)[:, :w].copy())
                planes.append(grp)
        k = self.L.ref_last_kcontrast()
        return pts.cpu().numpy().view(REF_POINT)[:n].copy(), planes, float(k)


    def detect_serialized(self, img_t, max_pts=100000, desc=True, dthreshold=0.001):
        """The reference pipeline with its sublevel merge made race-free WITHOUT touching its code.

        gCalcExtremaMap merges the sublevels of an octave with an unsynchronised check-then-write
        (akazed.cu:1364-1373, App. B-2); on a B200 the z-slices of a small octave are co-resident and the outcome varies.
        Here the stock scale space is built first (detect_keep), then hCalcExtremaMap is called once per sublevel on a
        copy of the octave's determinant planes in which every OTHER sublevel is zeroed (zero never passes the
        threshold), then the stock hNmsR, hRefine, hCalcOrient and hDescribe run on the same pyramid.
        Returns (points, planes, kcontrast)."""
        import torch
        L = self.L
        dev = img_t.device
        pts = torch.zeros(max_pts * 104, dtype=torch.uint8, device=dev)
        tmem = C.c_void_p()
        oparams = (C.c_int * 64)()
        noct = C.c_int()
        torch.cuda.synchronize()
        L.ref_akazer_detect_keep(self.hnd, C.c_void_p(img_t.data_ptr()), self.w, self.h, self.pitch, 0,
                                 C.c_void_p(pts.data_ptr()), max_pts, C.byref(tmem), oparams, C.byref(noct))
        torch.cuda.synchronize()
        no, S = noct.value, self.S
        osizes = list(oparams[0:no])
        offsets = list(oparams[no:2 * no + 1])
        owhps = [tuple(oparams[2 * no + 1 + 3 * k: 2 * no + 4 + 3 * k]) for k in range(no)]
        total = offsets[no]
        mem = _wrap_device(tmem.value, total * 4).view(torch.int32)
        W, H, P0 = owhps[0]
        msz = H * P0
        mem[0:2 * msz] = int(np.array([0xBDBDBDBD], dtype=np.uint32).view(np.int32)[0])     # akaze.cpp:253-258 (App. B-8)
        mem[2 * msz:3 * msz] = -1
        sched = ref_schedule(no, S)
        psz, neigh = 10000.0, 0
        f32 = mem.view(torch.float32)
        for o in range(no):
            w, h, p = owhps[o]
            sizes, ss, borders = sched[o]
            params = np.concatenate([borders, sizes]).astype(np.float32)
            psz = min(psz, float(borders[0]) * (1 << o))
            neigh = max(neigh, max(ss))
            dets = f32[offsets[o] + S * osizes[o]: offsets[o] + 2 * S * osizes[o]]           # group 1 = det (akaze.cpp:315-320)
            for j in range(S):
                scratch = torch.zeros_like(dets)
                scratch[j * osizes[o]:(j + 1) * osizes[o]] = dets[j * osizes[o]:(j + 1) * osizes[o]]
                torch.cuda.synchronize()
                L.ref_hCalcExtremaMap(C.c_void_p(scratch.data_ptr()), C.c_void_p(tmem.value), C.c_void_p(tmem.value + 4 * msz),
                                      C.c_void_p(tmem.value + 8 * msz), params.ctypes.data_as(C.POINTER(C.c_float)),
                                      o, S, dthreshold, w, h, p, P0)
        pts.zero_()
        L.ref_setMaxNumPoints(max_pts)
        L.ref_resetPointCounter()
        L.ref_hNmsR(C.c_void_p(pts.data_ptr()), C.c_void_p(tmem.value), C.c_void_p(tmem.value + 4 * msz), C.c_void_p(tmem.value + 8 * msz),
                    int(psz), neigh, W, H, P0)
        n = min(int(L.ref_readPointCounter()), max_pts)
        L.ref_hRefine(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S)
        if desc:
            L.ref_hCalcOrient(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S)
            L.ref_hDescribe(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S, 10)
        torch.cuda.synchronize()
        host = f32.cpu().numpy()
        planes = []
        for o in range(no):
            w, h, p = owhps[o]
            for sidx in range(S):
                grp = []
                for which in range(4):
                    off = offsets[o] + (which * S + sidx) * osizes[o]
                    grp.append(host[off:off + osizes[o]].reshape(h, p)[:, :w].copy())
                planes.append(grp)
        out = pts.cpu().numpy().view(REF_POINT)[:n].copy()
        del mem, f32
        L.ref_cuda_free(tmem)
        return out, planes, float(L.ref_last_kcontrast())

    def fast_detect_serialized(self, img_t, max_pts=100000, desc=True, ik_inject=0):
        """Integer pipeline (Akazer::fastDetect, akaze.cpp:506-743) with the sublevel merge serialised, as detect_serialized
        does for the float one.  img_t: (h, pitch) uint8 device tensor.  Returns (points, int planes, integer kcontrast)."""
        import torch
        L = self.L
        dev = img_t.device
        pts = torch.zeros(max_pts * 104, dtype=torch.uint8, device=dev)
        tmem = C.c_void_p()
        oparams = (C.c_int * 64)()
        noct = C.c_int()
        L.ref_fast_set_kcontrast_mode(1 if ik_inject > 0 else 0, int(ik_inject))
        torch.cuda.synchronize()
        L.ref_akazer_fast_detect_keep(self.hnd, C.c_void_p(img_t.data_ptr()), self.w, self.h, self.pitch, 0,
                                      C.c_void_p(pts.data_ptr()), max_pts, C.byref(tmem), oparams, C.byref(noct))
        torch.cuda.synchronize()
        L.ref_fast_set_kcontrast_mode(0, 0)
        no, S = noct.value, self.S
        osizes = list(oparams[0:no])
        offsets = list(oparams[no:2 * no + 1])
        owhps = [tuple(oparams[2 * no + 1 + 3 * k: 2 * no + 4 + 3 * k]) for k in range(no)]
        total = offsets[no]
        mem = _wrap_device(tmem.value, total * 4).view(torch.int32)
        W, H, P0 = owhps[0]
        msz = H * P0
        mem[0:msz] = int(np.array([0xC0C0C0C0], dtype=np.uint32).view(np.int32)[0])       # akaze.cpp:521 (memset byte of -1E6)
        mem[msz:2 * msz] = 0                                                                   # akaze.cpp:522 (low byte of the bits of -1e6f)
        mem[2 * msz:3 * msz] = -1
        sched = ref_schedule(no, S)
        psz, neigh = 10000.0, 0
        for o in range(no):
            w, h, p = owhps[o]
            sizes, ss, borders = sched[o]
            params = np.concatenate([borders, sizes]).astype(np.float32)
            psz = min(psz, float(borders[0]) * (1 << o))
            neigh = max(neigh, max(ss))
            dets = mem[offsets[o] + S * osizes[o]: offsets[o] + 2 * S * osizes[o]]
            for j in range(S):
                scratch = torch.zeros_like(dets)
                scratch[j * osizes[o]:(j + 1) * osizes[o]] = dets[j * osizes[o]:(j + 1) * osizes[o]]
                torch.cuda.synchronize()
                L.ref_fast_hCalcExtremaMap(C.c_void_p(scratch.data_ptr()), C.c_void_p(tmem.value), C.c_void_p(tmem.value + 4 * msz),
                                           C.c_void_p(tmem.value + 8 * msz), params.ctypes.data_as(C.POINTER(C.c_float)),
                                           o, S, 65, w, h, p, P0)
        pts.zero_()
        L.ref_setMaxNumPoints(max_pts)
        L.ref_resetPointCounter()
        L.ref_fast_hNmsR(C.c_void_p(pts.data_ptr()), C.c_void_p(tmem.value), C.c_void_p(tmem.value + 4 * msz), C.c_void_p(tmem.value + 8 * msz),
                         int(psz), neigh, W, H, P0)
        n = min(int(L.ref_readPointCounter()), max_pts)
        L.ref_fast_hRefine(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S)
        if desc:
            L.ref_fast_hCalcOrient(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S)
            L.ref_fast_hDescribe(C.c_void_p(pts.data_ptr()), n, max_pts, C.c_void_p(tmem.value), no, S, 10)
        torch.cuda.synchronize()
        host = mem.cpu().numpy()
        planes = []
        for o in range(no):
            w, h, p = owhps[o]
            for sidx in range(S):
                grp = []
                for which in range(4):
                    off = offsets[o] + (which * S + sidx) * osizes[o]
                    grp.append(host[off:off + osizes[o]].reshape(h, p)[:, :w].copy())
                planes.append(grp)
        out = pts.cpu().numpy().view(REF_POINT)[:n].copy()
        k = int(L.ref_fast_last_kcontrast())
        del mem
        L.ref_cuda_free(tmem)
        return out, planes, k


# ---- seeded inputs ------------------------------------------------------------------------------------------
def read_pgm(path):
    with open(path, "rb") as fh:
        b = fh.read()
    assert b[:2] == b"P5"
    toks, pos = [], 2
    while len(toks) < 3:
        while b[pos:pos + 1].isspace():
            pos += 1
        if b[pos:pos + 1] == b"#":
            while b[pos:pos + 1] != b"\n":
                pos += 1
            continue
        s = pos
        while not b[pos:pos + 1].isspace():
            pos += 1
        toks.append(int(b[s:pos]))
    pos += 1
    w, h, _ = toks
    return np.frombuffer(b, dtype=np.uint8, count=w * h, offset=pos).reshape(h, w).copy()


def synth_noise_u8(w, h, seed=0, sigma=2.0):
    """SURVEY 8d metric-1 input: uniform u8 noise, Gaussian blur sigma, min-max stretched to 0..255."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32)
    a = gaussian_filter(a, sigma, mode="reflect")
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-6)
    return np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)


def synth_shapes_u8(w, h, seed=1, nshapes=200):
    """'natural statistics' variant: random anti-aliased rectangles and discs, blurred with sigma 1."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    ss = 2
    img = np.full((h * ss, w * ss), 0.5, dtype=np.float32)
    yy, xx = np.mgrid[0:h * ss, 0:w * ss]
    for _ in range(nshapes):
        v = rng.uniform(0.0, 1.0)
        cx, cy = rng.uniform(0, w * ss), rng.uniform(0, h * ss)
        r = rng.uniform(4, min(w, h) * ss / 6)
        if rng.random() < 0.5:
            x0, x1 = int(max(cx - r, 0)), int(min(cx + r, w * ss))
            y0, y1 = int(max(cy - r * rng.uniform(0.3, 1.0), 0)), int(min(cy + r, h * ss))
            img[y0:y1, x0:x1] = v
        else:
            x0, x1 = int(max(cx - r, 0)), int(min(cx + r + 1, w * ss))
            y0, y1 = int(max(cy - r, 0)), int(min(cy + r + 1, h * ss))
            sub = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
            img[y0:y1, x0:x1][sub] = v
    img = img.reshape(h, ss, w, ss).mean(axis=(1, 3))
    img = gaussian_filter(img, 1.0, mode="reflect")
    img += rng.normal(0, 0.01, size=img.shape).astype(np.float32)
    return np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)


def u8_to_unit(a):
    """float(v) * float(1/255.) — what cv::Mat::convertTo(CV_32F, 1/255.) computes (main.cpp:149)."""
    return a.astype(np.float32) * np.float32(1.0 / 255.0)


def random_descriptors(n, seed=0):
    """SURVEY 8d metric-2 input: random 64-byte rows with 486 valid bits (bytes 61..63 and the top two bits of byte 60 zero)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    d[:, 61:] = 0
    d[:, 60] &= 0x3F
    return d
