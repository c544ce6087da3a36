"""ctypes binding of libakaze_b200.so (the C ABI in include/akaze_b200.h).

Host-side mirror used by tests/ and bench.py; torch supplies device memory and streams only.
There is no CPU fallback: if the CUDA library is missing, importing the symbols raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libakaze_b200.so")

AKZ_F32, AKZ_U8 = 0, 1
MATCH_COMPAT, MATCH_KNN2, MATCH_UNIQUE2 = 0, 1, 2
PLANE_LT, PLANE_DET, PLANE_LX, PLANE_LY = 0, 1, 2, 3


class Options(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("noctaves", C.c_int), ("max_scale", C.c_int),
                ("per", C.c_float), ("kcontrast", C.c_float), ("soffset", C.c_float), ("reordering", C.c_int),
                ("derivative_factor", C.c_float), ("dthreshold", C.c_float), ("diffusivity", C.c_int),
                ("descriptor_pattern_size", C.c_int), ("max_pts", C.c_int), ("max_batch", C.c_int),
                ("device", C.c_int), ("kcontrast_override", C.c_float), ("fused", C.c_int),
                ("fast_kcontrast_override", C.c_int), ("lanes", C.c_int)]


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("layer", "<i4"), ("ix", "<i4"), ("iy", "<i4")])
MATCH_DTYPE = np.dtype([("idx1", "<i4"), ("dist1", "<i4"), ("idx2", "<i4"), ("dist2", "<i4")])

# every symbol include/akaze_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "akz_version", "akz_last_error", "akz_default_options", "akz_fed_tau", "akz_fed_tau_internal", "akz_gauss_taps", "akz_compare_indices",
    "akz_create", "akz_destroy", "akz_sync", "akz_stream", "akz_num_levels", "akz_level_info", "akz_level_plane",
    "akz_launch_count", "akz_detect_and_compute", "akz_detect_and_compute_host", "akz_build_scale_space",
    "akz_get_kcontrast", "akz_lowpass", "akz_down_with_smooth", "akz_scharr_contrast", "akz_flow", "akz_nld_step",
    "akz_fed_cycle", "akz_hessian", "akz_match", "akz_match_merge", "akz_match_host", "akz_pack_points",
    "akz_unpack_desc", "akz_scatter_matches", "akz_orient", "akz_describe", "akz_detect_keypoints",
    "akz_profile_enable", "akz_profile_read", "akz_profile_octaves", "akz_profile_class_name", "akz_keypoints_to_opencv", "akz_matches_to_opencv",
    "akz_set_match_kernel", "akz_set_describe_kernel", "akz_plan_chunks", "akz_plan_match", "akz_fast_detect_and_compute", "akz_fast_detect_and_compute_host", "akz_fast_build_scale_space", "akz_fast_get_kcontrast", "akz_fast_lowpass",
    "akz_match_pairs", "akz_comm_unique_id", "akz_comm_init", "akz_comm_attach", "akz_comm_destroy", "akz_match_sharded",
    "akz_fast_down_with_smooth", "akz_fast_scharr_contrast", "akz_fast_flow", "akz_fast_nld_step", "akz_fast_hessian",
]
NUM_KCLASS = 13

_lib = None


def lib():
    """Load the CUDA library; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C cuda-akaze_b200` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    L.akz_last_error.restype = C.c_char_p
    L.akz_default_options.argtypes = [C.POINTER(Options)]
    L.akz_fed_tau.argtypes = [f, i, f, i, C.POINTER(C.c_float), i]
    L.akz_fed_tau_internal.argtypes = [i, f, f, i, C.POINTER(C.c_float), i]
    L.akz_gauss_taps.argtypes = [f, i, C.POINTER(C.c_float)]
    L.akz_compare_indices.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.akz_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
    L.akz_destroy.argtypes = [vp]
    L.akz_destroy.restype = None
    L.akz_sync.argtypes = [vp]
    L.akz_stream.argtypes = [vp]
    L.akz_stream.restype = vp
    L.akz_num_levels.argtypes = [vp]
    L.akz_launch_count.argtypes = [vp]
    L.akz_level_info.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(f),
                                 C.POINTER(i), C.POINTER(f), i]
    L.akz_level_plane.argtypes = [vp, i, i, i]
    L.akz_level_plane.restype = vp
    L.akz_detect_and_compute.argtypes = [vp, vp, i, i, i, i, i, ll, i, vp, vp, vp]
    L.akz_detect_and_compute_host.argtypes = [vp, vp, i, i, i, i, i, ll, i, vp, vp, vp]
    L.akz_build_scale_space.argtypes = [vp, vp, i, i, i, i, i, ll]
    L.akz_get_kcontrast.argtypes = [vp, C.POINTER(C.c_float), i]
    L.akz_lowpass.argtypes = [vp, vp, vp, i, i, i, ll, i, f, i]
    L.akz_down_with_smooth.argtypes = [vp, vp, vp, vp, i, i, i, ll, i, i, i, ll, i]
    L.akz_scharr_contrast.argtypes = [vp, vp, vp, f, i, i, i, ll, i]
    L.akz_flow.argtypes = [vp, vp, vp, i, vp, f, i, i, i, ll, i]
    L.akz_nld_step.argtypes = [vp, vp, vp, vp, f, i, i, i, ll, i]
    L.akz_fed_cycle.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_float), i, i, i, i, ll, i]
    L.akz_hessian.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, ll, i]
    L.akz_orient.argtypes = [vp, vp, vp, i]
    L.akz_describe.argtypes = [vp, vp, vp, vp, i]
    L.akz_detect_keypoints.argtypes = [vp, i, vp, vp]
    L.akz_match.argtypes = [vp, vp, i, vp, i, i, i, i, vp]
    L.akz_match_merge.argtypes = [vp, vp, i, i, i, i, vp]
    L.akz_match_host.argtypes = [vp, vp, i, vp, i, i, vp]
    L.akz_match_pairs.argtypes = [vp, vp, vp, i, i, vp]
    L.akz_comm_unique_id.argtypes = [vp]
    L.akz_comm_init.argtypes = [vp, i, i, vp]
    L.akz_comm_attach.argtypes = [vp, vp, i, i]
    L.akz_comm_destroy.argtypes = [vp]
    L.akz_match_sharded.argtypes = [vp, vp, i, vp, i, i, i, vp]
    L.akz_pack_points.argtypes = [vp, vp, vp, vp, vp, i, i]
    L.akz_unpack_desc.argtypes = [vp, vp, i, vp]
    L.akz_scatter_matches.argtypes = [vp, vp, i, vp, vp]
    L.akz_fast_detect_and_compute.argtypes = [vp, vp, i, i, i, i, ll, i, vp, vp, vp]
    L.akz_fast_detect_and_compute_host.argtypes = [vp, vp, i, i, i, i, ll, i, vp, vp, vp]
    L.akz_fast_build_scale_space.argtypes = [vp, vp, i, i, i, i, ll]
    L.akz_fast_get_kcontrast.argtypes = [vp, C.POINTER(C.c_int), i]
    L.akz_fast_lowpass.argtypes = [vp, vp, i, vp, vp, i, i, i, ll, i, f, i]
    L.akz_fast_down_with_smooth.argtypes = [vp, vp, vp, vp, i, i, i, ll, i, i, i, ll, i]
    L.akz_fast_scharr_contrast.argtypes = [vp, vp, vp, vp, f, i, i, i, ll, i]
    L.akz_fast_flow.argtypes = [vp, vp, vp, i, vp, i, i, i, ll, i]
    L.akz_fast_nld_step.argtypes = [vp, vp, vp, vp, f, i, i, i, ll, i]
    L.akz_fast_hessian.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, ll, i]
    L.akz_set_match_kernel.argtypes = [i]
    L.akz_plan_chunks.argtypes = [i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_int), i]
    L.akz_plan_match.argtypes = [i, i, C.POINTER(C.c_int)]
    L.akz_set_match_kernel.restype = None
    L.akz_set_describe_kernel.argtypes = [i]
    L.akz_set_describe_kernel.restype = None
    L.akz_keypoints_to_opencv.argtypes = [vp, i, i, vp]
    L.akz_matches_to_opencv.argtypes = [vp, i, vp]
    L.akz_profile_enable.argtypes = [vp, i]
    L.akz_profile_read.argtypes = [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.akz_profile_octaves.argtypes = [vp, i, i, C.POINTER(C.c_double)]
    L.akz_profile_class_name.argtypes = [i]
    L.akz_profile_class_name.restype = C.c_char_p
    _lib = L
    return L


class AkazeError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise AkazeError(f"akz error {rc}: {lib().akz_last_error().decode()}")


def default_options(**kw):
    o = Options()
    lib().akz_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def fed_tau(T, M=1, tau_max=0.25, reordering=True):
    buf = (C.c_float * 1024)()
    n = lib().akz_fed_tau(T, M, tau_max, int(reordering), buf, 1024)
    return np.array(buf[:max(n, 0)], dtype=np.float32)


def gauss_taps(var, radius):
    buf = (C.c_float * (radius + 1))()
    lib().akz_gauss_taps(var, radius, buf)
    return np.array(buf[:], dtype=np.float32)


def compare_indices():
    a, b = (C.c_int * 488)(), (C.c_int * 488)()
    lib().akz_compare_indices(a, b)
    return np.array(a[:486], dtype=np.int32), np.array(b[:486], dtype=np.int32)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Context:
    """One akz_ctx: owns the pyramid buffers of `max_batch` frames of a fixed size on one device."""

    def __init__(self, width=0, height=0, **kw):
        import torch
        self.torch = torch
        self.opt = default_options(width=width, height=height, **kw)
        h = C.c_void_p()
        _check(lib().akz_create(C.byref(self.opt), C.byref(h)))
        self.h = h
        self.device = torch.device("cuda", torch.cuda.current_device() if self.opt.device < 0 else self.opt.device)

    def close(self):
        if getattr(self, "h", None):
            lib().akz_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(lib().akz_sync(self.h))

    @property
    def stream_handle(self):
        return lib().akz_stream(self.h)

    def torch_stream(self):
        return self.torch.cuda.ExternalStream(self.stream_handle, device=self.device)

    @property
    def launches(self):
        return lib().akz_launch_count(self.h)

    def profile(self, on):
        _check(lib().akz_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{class name: (milliseconds, launches)} accumulated since the last read (synchronises)."""
        ms = (C.c_double * NUM_KCLASS)()
        n = (C.c_longlong * NUM_KCLASS)()
        _check(lib().akz_profile_read(self.h, NUM_KCLASS, ms, n))
        return {lib().akz_profile_class_name(k).decode(): (ms[k], n[k]) for k in range(NUM_KCLASS) if n[k]}

    def profile_octaves(self, noct=8):
        """{class name: [ms of octave 0, 1, ...]} of the last profile_read."""
        ms = (C.c_double * (NUM_KCLASS * noct))()
        _check(lib().akz_profile_octaves(self.h, NUM_KCLASS, noct, ms))
        out = {}
        for k in range(NUM_KCLASS):
            row = [ms[k * noct + o] for o in range(noct)]
            if any(row):
                out[lib().akz_profile_class_name(k).decode()] = row
        return out

    @property
    def num_levels(self):
        return lib().akz_num_levels(self.h)

    def level_info(self, level):
        w, h, p, n, ss = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        sz = C.c_float()
        tau = (C.c_float * 128)()
        _check(lib().akz_level_info(self.h, level, C.byref(w), C.byref(h), C.byref(p), C.byref(n), C.byref(sz), C.byref(ss), tau, 128))
        return dict(w=w.value, h=h.value, pitch=p.value, nsteps=n.value, size=sz.value, sigma_size=ss.value,
                    tau=np.array(tau[:n.value], dtype=np.float32))

    def plane(self, level, which, frame=0):
        """Copy one plane of the last processed chunk to the host as an (h, w) float32 array."""
        info = self.level_info(level)
        ptr = lib().akz_level_plane(self.h, level, which, frame)
        if not ptr:
            raise AkazeError("plane not available")
        torch = self.torch
        self.sync()
        n = info["pitch"] * info["h"]
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        # device-to-device copy through the CUDA runtime that torch already loaded
        _cudart_memcpy_d2d(out.data_ptr(), ptr, n * 4)
        return out.cpu().numpy().reshape(info["h"], info["pitch"])[:, :info["w"]].copy()

    def kcontrast(self, nframes=1):
        buf = (C.c_float * nframes)()
        _check(lib().akz_get_kcontrast(self.h, buf, nframes))
        return np.array(buf[:], dtype=np.float32)

    # ---- hot path ------------------------------------------------------------------------------
    def _img_args(self, images):
        torch = self.torch
        assert images.is_cuda and images.dim() == 3 and images.is_contiguous()
        n, h, pitch = images.shape
        dtype = AKZ_U8 if images.dtype == torch.uint8 else AKZ_F32
        if dtype == AKZ_F32:
            assert images.dtype == torch.float32
        return n, h, pitch, dtype

    def build_scale_space(self, images, width=None):
        n, h, pitch, dtype = self._img_args(images)
        w = width or self.opt.width
        _check(lib().akz_build_scale_space(self.h, _ptr(images), dtype, n, w, h, pitch, pitch * h))

    def alloc_results(self, nframes, describe=True):
        torch = self.torch
        mp = self.opt.max_pts
        counts = torch.zeros(nframes, dtype=torch.int32, device=self.device)
        kpts = torch.zeros(nframes, mp, 8, dtype=torch.int32, device=self.device)
        desc = torch.zeros(nframes, mp, 64, dtype=torch.uint8, device=self.device) if describe else None
        return counts, kpts, desc

    def detect_and_compute(self, images, describe=True, out=None, width=None):
        """images: (n, h, pitch) float32 in [0,1] or uint8, on the device.  Returns device tensors
        (counts[n], kpts[n, max_pts, 8 x int32 words = akz_keypoint], desc[n, max_pts, 64])."""
        n, h, pitch, dtype = self._img_args(images)
        w = width or self.opt.width
        counts, kpts, desc = out if out is not None else self.alloc_results(n, describe)
        _check(lib().akz_detect_and_compute(self.h, _ptr(images), dtype, n, w, h, pitch, pitch * h, int(describe),
                                            _ptr(counts), _ptr(kpts), _ptr(desc)))
        return counts, kpts, desc

    # ---- integer ("fast") pipeline -------------------------------------------------------------------
    def fast_detect_and_compute(self, images, describe=True, out=None, width=None):
        """images: (n, h, pitch) uint8 on the device -> (counts, kpts, desc) as detect_and_compute, integer arithmetic."""
        n, h, pitch, dtype = self._img_args(images)
        assert dtype == AKZ_U8
        w = width or self.opt.width
        counts, kpts, desc = out if out is not None else self.alloc_results(n, describe)
        _check(lib().akz_fast_detect_and_compute(self.h, _ptr(images), n, w, h, pitch, pitch * h, int(describe),
                                                 _ptr(counts), _ptr(kpts), _ptr(desc)))
        return counts, kpts, desc

    def fast_build_scale_space(self, images, width=None):
        n, h, pitch, dtype = self._img_args(images)
        assert dtype == AKZ_U8
        _check(lib().akz_fast_build_scale_space(self.h, _ptr(images), n, width or self.opt.width, h, pitch, pitch * h))

    def fast_kcontrast(self, nframes=1):
        buf = (C.c_int * nframes)()
        _check(lib().akz_fast_get_kcontrast(self.h, buf, nframes))
        return np.array(buf[:], dtype=np.int32)

    def plane_int(self, level, which, frame=0):
        """A plane of the integer pipeline as int32 (the buffers are shared with the float pipeline)."""
        return self.plane(level, which, frame).view(np.int32)

    def detect_and_compute_host(self, images, describe=True, out=None, width=None, fast=False):
        """images: (n, h, pitch) numpy array (or pinned torch CPU tensor).  Returns numpy arrays.
        fast=True: the integer pipeline (uint8 frames only)."""
        arr = images
        n, h, pitch = arr.shape
        is_u8 = str(arr.dtype).endswith("uint8")
        dtype = AKZ_U8 if is_u8 else AKZ_F32
        w = width or self.opt.width
        mp = self.opt.max_pts
        if out is None:
            counts = np.zeros(n, dtype=np.int32)
            kpts = np.zeros((n, mp), dtype=KEYPOINT_DTYPE)
            desc = np.zeros((n, mp, 64), dtype=np.uint8) if describe else None
        else:
            counts, kpts, desc = out
        ip = arr.data_ptr() if hasattr(arr, "data_ptr") else arr.ctypes.data
        outp = (C.c_void_p(_host_ptr(counts)), C.c_void_p(_host_ptr(kpts)), C.c_void_p(_host_ptr(desc)) if desc is not None else C.c_void_p(0))
        if fast:
            assert is_u8
            _check(lib().akz_fast_detect_and_compute_host(self.h, C.c_void_p(ip), n, w, h, pitch, pitch * h, int(describe), *outp))
        else:
            _check(lib().akz_detect_and_compute_host(self.h, C.c_void_p(ip), dtype, n, w, h, pitch, pitch * h, int(describe), *outp))
        return counts, kpts, desc

    # ---- stage seams -----------------------------------------------------------------------------
    def lowpass(self, src, dst, w, var, ksz):
        n, h, p = src.shape
        _check(lib().akz_lowpass(self.h, _ptr(src), _ptr(dst), w, h, p, p * h, n, var, ksz))

    def down_with_smooth(self, src, sw, dst, smooth, dw):
        n, sh, sp = src.shape
        _, dh, dp = dst.shape
        _check(lib().akz_down_with_smooth(self.h, _ptr(src), _ptr(dst), _ptr(smooth), sw, sh, sp, sp * sh, dw, dh, dp, dp * dh, n))

    def scharr_contrast(self, src, w, per=0.7):
        torch = self.torch
        n, h, p = src.shape
        k = torch.zeros(n, dtype=torch.float32, device=self.device)
        _check(lib().akz_scharr_contrast(self.h, _ptr(src), _ptr(k), per, w, h, p, p * h, n))
        self.sync()
        return k

    def flow(self, src, flow, w, k, type=1, kscale=1.0):
        n, h, p = src.shape
        _check(lib().akz_flow(self.h, _ptr(src), _ptr(flow), type, _ptr(k), kscale, w, h, p, p * h, n))

    def nld_step(self, src, flow, dst, w, tau):
        n, h, p = src.shape
        _check(lib().akz_nld_step(self.h, _ptr(src), _ptr(flow), _ptr(dst), tau, w, h, p, p * h, n))

    def fed_cycle(self, src, flow, dst, tmp, w, tau):
        n, h, p = src.shape
        t = (C.c_float * len(tau))(*[float(x) for x in tau])
        _check(lib().akz_fed_cycle(self.h, _ptr(src), _ptr(flow), _ptr(dst), _ptr(tmp), t, len(tau), w, h, p, p * h, n))

    def hessian(self, smooth, lx, ly, det, w, step):
        n, h, p = smooth.shape
        _check(lib().akz_hessian(self.h, _ptr(smooth), _ptr(lx), _ptr(ly), _ptr(det), step, w, h, p, p * h, n))

    def orient(self, counts, kpts):
        _check(lib().akz_orient(self.h, _ptr(counts), _ptr(kpts), counts.shape[0]))

    def describe(self, counts, kpts, desc):
        _check(lib().akz_describe(self.h, _ptr(counts), _ptr(kpts), _ptr(desc), counts.shape[0]))

    def detect_keypoints(self, nframes):
        counts, kpts, _ = self.alloc_results(nframes, describe=False)
        _check(lib().akz_detect_keypoints(self.h, nframes, _ptr(counts), _ptr(kpts)))
        return counts, kpts

    # ---- matcher ------------------------------------------------------------------------------------
    def match(self, q, t, mode=MATCH_COMPAT, t_index_base=0, finalize=True, out=None):
        """q: (nq, 64) uint8, t: (nt, 64) uint8 device tensors -> (nq, 4) int32 device tensor."""
        torch = self.torch
        nq, nt = q.shape[0], t.shape[0]
        res = out if out is not None else torch.zeros(nq, 4, dtype=torch.int32, device=self.device)
        _check(lib().akz_match(self.h, _ptr(q), nq, _ptr(t), nt, t_index_base, mode, int(finalize), _ptr(res)))
        return res

    def match_merge(self, parts, mode, finalize=True):
        torch = self.torch
        nparts, nq = parts.shape[0], parts.shape[1]
        res = torch.zeros(nq, 4, dtype=torch.int32, device=self.device)
        _check(lib().akz_match_merge(self.h, _ptr(parts), nparts, nq, mode, int(finalize), _ptr(res)))
        return res

    def match_pairs(self, desc, counts, mode=MATCH_COMPAT, out=None):
        """desc (nf, max_pts, 64), counts (nf,) as detect_and_compute returned them -> (nf, max_pts, 4) int32: frame f matched
        against frame f - 1 (row 0 unused); counts are read on the device."""
        torch = self.torch
        nf = desc.shape[0]
        res = out if out is not None else torch.zeros(nf, self.opt.max_pts, 4, dtype=torch.int32, device=self.device)
        _check(lib().akz_match_pairs(self.h, _ptr(desc), _ptr(counts), nf, mode, _ptr(res)))
        return res

    # ---- train-sharded matching (one process per GPU): the gather runs inside the library on its stream ----
    def comm_init(self, nranks, rank, unique_id):
        """unique_id: the 128 bytes akaze_b200.comm_unique_id() produced on one rank, handed to all ranks by the caller."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(lib().akz_comm_init(self.h, nranks, rank, buf))

    def comm_destroy(self):
        _check(lib().akz_comm_destroy(self.h))

    def match_sharded(self, q, t_local, t_index_base, mode=MATCH_KNN2, out=None):
        """q: (nq, 64) all queries, t_local: this rank's contiguous train range starting at global index t_index_base.
        akz_match(finalize=0) -> ncclAllGather -> akz_match_merge(finalize=1), all on the context's stream; no host sync."""
        torch = self.torch
        nq = q.shape[0]
        res = out if out is not None else torch.zeros(nq, 4, dtype=torch.int32, device=self.device)
        _check(lib().akz_match_sharded(self.h, _ptr(q), nq, _ptr(t_local), t_local.shape[0], t_index_base, mode, _ptr(res)))
        return res

    def match_host(self, q, t, mode=MATCH_COMPAT):
        nq, nt = q.shape[0], t.shape[0]
        res = np.zeros((nq, 4), dtype=np.int32)
        _check(lib().akz_match_host(self.h, C.c_void_p(q.ctypes.data), nq, C.c_void_p(t.ctypes.data), nt, mode,
                                    C.c_void_p(res.ctypes.data)))
        return res


def comm_unique_id():
    """128-byte NCCL id for Context.comm_init (create on one rank, distribute to the others)."""
    buf = (C.c_char * 128)()
    _check(lib().akz_comm_unique_id(buf))
    return bytes(buf.raw)


def _host_ptr(a):
    if a is None:
        return 0
    return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data


_cudart = None


def _cudart_memcpy_d2d(dst, src, nbytes):
    """cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice) via torch's tensor machinery."""
    import torch
    # wrap the foreign pointer without taking ownership
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(src), False), "version": 2}
    srct = torch.as_tensor(h, device="cuda")
    dstt_h = _Holder()
    dstt_h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(dst), False), "version": 2}
    dstt = torch.as_tensor(dstt_h, device="cuda")
    dstt.copy_(srct)
    torch.cuda.synchronize()


def keypoints_to_opencv(kpts, max_scale=4):
    """KEYPOINT_DTYPE array (host) -> (n, 7) float32: pt.x, pt.y, size, angle (deg), response, octave, class_id."""
    k = np.ascontiguousarray(kpts)
    out = np.zeros((len(k), 7), dtype=np.float32)
    _check(lib().akz_keypoints_to_opencv(C.c_void_p(k.ctypes.data), len(k), max_scale, C.c_void_p(out.ctypes.data)))
    return out


def matches_to_opencv(m):
    """(nq, 4) int32 match results (host) -> (n_accepted, 3) int32: queryIdx, trainIdx, distance."""
    a = np.ascontiguousarray(m, dtype=np.int32)
    out = np.zeros((len(a), 3), dtype=np.int32)
    n = lib().akz_matches_to_opencv(C.c_void_p(a.ctypes.data), len(a), C.c_void_p(out.ctypes.data))
    if n < 0:
        _check(n)
    return out[:n]


def keypoints_from_words(words):
    """(…, 8) int32 array as returned by detect_and_compute -> structured KEYPOINT_DTYPE array."""
    a = np.ascontiguousarray(words)
    return a.view(KEYPOINT_DTYPE).reshape(a.shape[:-1])
