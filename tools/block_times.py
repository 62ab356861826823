#!/usr/bin/env python
"""When do the simulation kernel's blocks finish?  (at_debug_block_times; development tooling.)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200 import MultiAgentTeamEnv, _capi  # noqa: E402
from alphatank_b200.configs.config_teams import team_vs_bot_configs  # noqa: E402

dev = torch.device("cuda:0")
E_ = 65536
env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=E_, device=dev, seed=0)
core = env.core
env.reset()
acts = torch.empty((64, E_, 2, 3), dtype=torch.int32, device=dev)
for t in range(64):
    core.synthetic_actions(t, acts[t])
for t in range(600):
    core.step(acts[t % 64])
nb = (E_ + 63) // 64
times = torch.zeros(2 * nb, dtype=torch.int64, device=dev)
_capi.check(core.L.at_debug_block_times(core.h, C.c_void_p(times.data_ptr())), "dbg")
for rep in range(3):
    core.step(acts[rep])
    torch.cuda.synchronize()
    t = times.cpu().numpy().reshape(nb, 2).astype(np.int64)
    t0 = t[:, 0].min()
    end = np.sort(t[:, 1] - t0) / 1e3
    start = (t[:, 0] - t0) / 1e3
    print("rep %d: blocks start within %.1f us; finish: min %.1f  p10 %.1f  p25 %.1f  p50 %.1f  p75 %.1f  p90 %.1f  p99 %.1f  max %.1f us" % (
        rep, start.max(), end[0], *[end[int(q * (nb - 1))] for q in (0.10, 0.25, 0.5, 0.75, 0.9, 0.99)], end[-1]))
_capi.check(core.L.at_debug_block_times(core.h, None), "dbg")
