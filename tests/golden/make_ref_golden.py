"""Generates tests/golden/ref_320x240.npz from the UNMODIFIED reference (oracle/_ref/libref_akaze.so) on a GPU box.

The reference is a CUDA library, so this runs under gpurun:
    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_320x240.npz'
and the file is then copied to tests/golden/.  It pins the CPU oracle (tests/test_cpu.py::
test_oracle_against_reference_golden_fixture): sha256 of every Lt/det/Lx/Ly plane of the reference's pyramid plus
an 8x-subsampled copy of each, the reference's keypoints, angles and descriptors, and the contrast factor the
reference's (racy, App. B-1) reduction produced in that run.  The keypoints come from the reference's own kernels with
the sublevel merge serialised (bindings.RefAkazer.detect_serialized: the stock merge is a data race, App. B-2).  320x240 with 2 octaves is a size at which the
reference's blur kernels have no uninitialised halo rows (App. B-7).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cuda-akaze_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out):
    import torch
    import bindings as B
    seed, w, h = 7, 320, 240
    img = B.u8_to_unit(B.synth_shapes_u8(w, h, seed=seed))
    pitch = (w + 127) // 128 * 128
    buf = np.zeros((h, pitch), dtype=np.float32)
    buf[:, :w] = img
    d = torch.from_numpy(buf).cuda()
    ref = B.RefAkazer(w, h, pitch, noctaves=2)
    pts, planes, k = ref.detect_serialized(d, max_pts=20000, desc=True)
    g = {"seed": seed, "img_sha256": hashlib.sha256(img.tobytes()).hexdigest(), "kcontrast": np.float32(k), "nlevels": len(planes)}
    for l, grp in enumerate(planes):
        for which, nm in enumerate(("lt", "det", "lx", "ly")):
            a = np.ascontiguousarray(grp[which])
            g[f"sha_{l}_{nm}"] = hashlib.sha256(a.tobytes()).hexdigest()
            g[f"sub_{l}_{nm}"] = a[::8, ::8].copy()
    g["kp_x"], g["kp_y"], g["kp_layer"] = pts["x"].copy(), pts["y"].copy(), pts["octave"].copy()
    g["kp_size"], g["kp_angle"], g["kp_desc"] = pts["size"].copy(), pts["angle"].copy(), pts["features"].copy()
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(pts)} keypoints, k = {k:.6f}, {len(planes)} levels")
    ref.close()
    # integer pipeline (Akazer::fastDetect): int planes, keypoints of the serialised detector, integer contrast factor
    img8 = B.synth_shapes_u8(w, h, seed=seed)
    b8 = np.zeros((h, pitch), dtype=np.uint8)
    b8[:, :w] = img8
    ref = B.RefAkazer(w, h, pitch, noctaves=2)
    pts, planes, ik = ref.fast_detect_serialized(torch.from_numpy(b8).cuda(), max_pts=20000, desc=True)
    g = {"seed": seed, "img_sha256": hashlib.sha256(img8.tobytes()).hexdigest(), "kcontrast": np.int32(ik), "nlevels": len(planes)}
    for l, grp in enumerate(planes):
        for which, nm in enumerate(("lt", "det", "lx", "ly")):
            a = np.ascontiguousarray(grp[which]).astype(np.int32)
            g[f"sha_{l}_{nm}"] = hashlib.sha256(a.tobytes()).hexdigest()
            g[f"sub_{l}_{nm}"] = a[::8, ::8].copy()
    g["kp_x"], g["kp_y"], g["kp_layer"], g["kp_size"] = pts["x"].copy(), pts["y"].copy(), pts["octave"].copy(), pts["size"].copy()
    g["kp_angle"], g["kp_desc"] = pts["angle"].copy(), pts["features"].copy()
    out2 = out.replace("ref_320x240", "ref_fast_320x240")
    np.savez_compressed(out2, **g)
    print(f"wrote {out2}: {len(pts)} keypoints, integer k = {ik}, {len(planes)} levels")
    ref.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_320x240.npz"))
