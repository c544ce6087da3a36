"""H2D rate of this box: torch pinned memory vs cudaHostAlloc(WriteCombined), one big copy vs 32-frame pieces."""
import ctypes as C, time, torch, numpy as np
rt = C.CDLL("libcudart.so.12") if True else None
N = 256 * 1920 * 1080 * 4
dst = torch.empty(N, dtype=torch.uint8, device="cuda")
def rate(src_ptr, pieces, label):
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step = N // pieces
    best = 0
    for it in range(4):
        with torch.cuda.stream(s):
            e0.record(s)
            for p in range(pieces):
                rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr() + p * step), C.c_void_p(src_ptr + p * step), C.c_size_t(step), 1, C.c_void_p(s.cuda_stream))
            e1.record(s)
        e1.synchronize()
        best = max(best, N / e0.elapsed_time(e1) / 1e6)
    print(f"{label:50s} {best:6.2f} GB/s")
pin = torch.empty(N, dtype=torch.uint8).pin_memory()
rate(pin.data_ptr(), 1, "torch pinned, one copy")
rate(pin.data_ptr(), 8, "torch pinned, 8 pieces")
for flags, name in ((0, "cudaHostAlloc default"), (4, "cudaHostAlloc write-combined"), (1, "cudaHostAlloc portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), flags) == 0
    rate(p.value, 1, name + ", one copy")
    rate(p.value, 8, name + ", 8 pieces")
    rt.cudaFreeHost(p)
