#!/usr/bin/env python
"""Times the observation-rows variants of one tick at bench size: at_step (raw fp32 rows), at_step_norm (normalised fp32
rows only / raw + normalised) and at_step_norm16 (normalised fp16 rows), minus the simulation alone.  Development tooling."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200 import MultiAgentTeamEnv, _capi  # noqa: E402
from alphatank_b200.configs.config_teams import team_vs_bot_configs  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=N, device=dev)
core = env.core
env.reset()
acts = torch.empty((64, N, core.A, 3), dtype=torch.int32, device=dev)
for t in range(64):
    core.synthetic_actions(t, acts[t])
RD = core.R * core.D
raw = [torch.zeros((N, RD), device=dev) for _ in range(2)]
nrm = [torch.zeros((N, RD), device=dev) for _ in range(2)]
h16 = [torch.zeros((N, RD), device=dev, dtype=torch.float16) for _ in range(2)]


def sim_only(t):
    a = acts[t % 64]
    _capi.check(core.L.at_step(core.h, C.c_void_p(a.data_ptr()), None, C.c_void_p(core.rewards.data_ptr()), C.c_void_p(core.done.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "at_step")


variants = [
    ("simulation alone", sim_only),
    ("at_step: raw fp32 rows", lambda t: core.step(acts[t % 64], obs_out=raw[t % 2])),
    ("at_step_norm: normalised fp32 rows only", lambda t: core.step_norm(acts[t % 64], nrm[t % 2])),
    ("at_step_norm: raw + normalised fp32 rows", lambda t: core.step_norm(acts[t % 64], nrm[t % 2], obs_out=raw[t % 2])),
    ("at_step_norm16: normalised fp16 rows only", lambda t: core.step_norm16(acts[t % 64], h16[t % 2])),
    ("at_step_norm16: raw fp32 + normalised fp16 rows", lambda t: core.step_norm16(acts[t % 64], h16[t % 2], obs_out=raw[t % 2])),
]
for t in range(400):
    core.step(acts[t % 64])
for name, fn in variants[1:]:     # every buffer written in full once
    fn(0); fn(1)
base = None
variants = variants + [("PERSISTENT BUFFERS (at_set_rows_mode)", None)] + [(n, f) for n, f in variants[1:]]
for name, fn in variants:
    if fn is None:
        core.set_rows_mode(True)
        print(name, flush=True)
        continue
    for t in range(20):
        fn(t)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for t in range(200):
        fn(t)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 200
    if base is None:
        base = ms
    print("%-50s %.4f ms/tick  (rows: %.4f ms)" % (name, ms, ms - base), flush=True)
