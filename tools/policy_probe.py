#!/usr/bin/env python
"""Times at_policy_forward (csrc/at_policy_tc.cuh) alone: two policies x N rows of 528 columns, all 17 K chunks and with
the static block folded (9 chunks).  CUDA events, inputs larger than L2 (2 x N x 528 x 4 B).  Development tooling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200.learner.ppo_utils import FusedPolicies, PPOAgentPPO  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D, A = 528, 2
torch.manual_seed(0)
agents = [PPOAgentPPO(D).to(dev) for _ in range(A)]
obs = [torch.randn((N, A, D), device=dev) for _ in range(3)]       # rotate: 3 x 277 MB > L2
acts = torch.zeros((N, A, 3), dtype=torch.int32, device=dev)
lps, vals = torch.zeros((N, A, 3), device=dev), torch.zeros((N, A), device=dev)
ctr = torch.tensor([1, 0], dtype=torch.int64, device=dev)
for half, fold in ((False, False), (False, True), (True, False), (True, True)):
    fp = FusedPolicies(agents, D, dev, half=half)
    if half and obs[0].dtype != torch.float16:
        obs = [o.half() for o in obs] + [torch.randn((N, A, D), device=dev).half() for _ in range(3)]   # 6 x 138 MB > L2
    if fold:
        fp.set_static(237, 518, obs[0][0])
    for i in range(5):
        fp.forward(obs[i % len(obs)], acts, lps, vals, ctr)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 30
    ev0.record()
    for i in range(reps):
        fp.forward(obs[i % len(obs)], acts, lps, vals, ctr)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    k = bin(fp.chunk_mask).count("1") * fp.kc if fp.chunk_mask else (D + fp.kc - 1) // fp.kc * fp.kc
    flop = 2.0 * N * A * (k * 512 + 4 * 128 * 128 + 9 * 128)
    print("at_policy_forward  %d rows x %d policies  K = %3d  %.4f ms  %.1f TFLOP/s (%s)  %.1f M rows/s" % (
        N, A, k, ms, flop / ms / 1e9, "fp16" if half else "tf32", N * A / ms / 1e3), flush=True)
