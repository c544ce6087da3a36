#!/usr/bin/env python
"""L2 bandwidth reference for the profiles: copy between two 24 MiB buffers (both resident in the 126 MB L2) many times,
timed with CUDA events; read + write bytes / time.  torch is plumbing here (a plain copy kernel), not a product path."""
import json
import torch
n = 24 * 1024 * 1024 // 4
a = torch.rand(n, device="cuda"); b = torch.empty_like(a)
for _ in range(20):
    b.copy_(a)
torch.cuda.synchronize()
best = 0.0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        b.copy_(a)
    e1.record(); e1.synchronize()
    best = max(best, 2 * n * 4 * 200 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
big = torch.rand(1 << 28, device="cuda"); big2 = torch.empty_like(big)
for _ in range(3):
    big2.copy_(big)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    big2.copy_(big)
e1.record(); e1.synchronize()
hbm = 2 * big.numel() * 4 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps({"l2_resident_copy_gbs": round(best, 1), "hbm_copy_gbs": round(hbm, 1), "how": "torch copy_, 24 MiB buffers x200 (L2) / 1 GiB buffers x10 (HBM), read+write bytes"}))
