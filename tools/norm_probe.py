#!/usr/bin/env python
"""Stand-alone timing of at_tick_normalize over rotating buffers (cold in L2).  Development tooling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200.learner.ppo_utils import tick_normalize_cuda  # noqa: E402

dev = torch.device("cuda:0")
N, A, D, K = 65536, 2, 528, 6
xs = [torch.randn(N, A, D, device=dev) for _ in range(K)]
ys = [torch.empty(N, A, D, device=dev) for _ in range(K)]
for mode in ("out_of_place", "in_place", "copy"):
    for w in range(2):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for i in range(4 * K):
            if mode == "out_of_place":
                tick_normalize_cuda(xs[i % K], out=ys[i % K])
            elif mode == "in_place":
                tick_normalize_cuda(xs[i % K], out=xs[i % K])
            else:
                ys[i % K].copy_(xs[i % K])
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / (4 * K)
    print("%-13s %.1f us  %.2f TB/s (read + write)" % (mode, ms * 1e3, 2 * N * A * D * 4 / ms / 1e9))
