#!/usr/bin/env python
"""Pure-write HBM bandwidth of this GPU (torch fill_ / zero_ / a copy for comparison): the rows kernel is a write stream,
MEASURED_PEAKS.json's hbm_gbs is a copy (read + write)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29     # 2 GiB of floats
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for name, fn, nbytes in (("fill_", lambda: x.fill_(1.5), 4 * n), ("zero_ (memset)", lambda: x.zero_(), 4 * n),
                         ("copy_ (read + write)", lambda: y.copy_(x), 8 * n), ("sum (read)", lambda: x.sum(), 4 * n)):
    ms = t(fn)
    print("%-22s %.3f ms  %.1f GB/s" % (name, ms, nbytes / ms / 1e6))
