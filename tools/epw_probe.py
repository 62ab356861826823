#!/usr/bin/env python
"""Small-batch tick latency against the simulation kernel's envs-per-warp spread (AT_SIM_EPW, read at at_create).
Sim alone (at_step without an observation buffer) and the whole tick.  Development tooling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402
from alphatank_b200 import _capi  # noqa: E402
from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv  # noqa: E402
from alphatank_b200.configs.config_teams import team_vs_bot_configs  # noqa: E402

dev = torch.device("cuda:0")


def sim_only(core, a):
    _capi.check(core.L.at_step(core.h, C.c_void_p(a.data_ptr()), None, C.c_void_p(core.rewards.data_ptr()), C.c_void_p(core.done.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "at_step")


def run(name, make, steps=200, burn=300):
    env = make()
    core = env.core
    env.reset()
    A = core.A
    acts = torch.empty((64, core.num_envs, A, 3), dtype=torch.int32, device=dev)
    for t in range(64):
        core.synthetic_actions(t, acts[t])
    out = []
    for obs in (True, False):
        fn = (lambda t: core.step(acts[t % 64])) if obs else (lambda t: sim_only(core, acts[t % 64]))
        for t in range(burn):
            fn(t)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for t in range(steps):
            fn(t)
        ev1.record()
        torch.cuda.synchronize()
        out.append(ev0.elapsed_time(ev1) / steps)
    print("%-44s epw %-4s %7d envs  tick %.4f ms  sim %.4f ms" % (name, os.environ.get("AT_SIM_EPW", "auto"), core.num_envs, out[0], out[1]), flush=True)


cases = [
    ("2v2 team_vs_bot", lambda n: MultiAgentTeamEnv(team_vs_bot_configs, num_envs=n, device=dev)),
    ("config 2: 1v1 bot_agent vs smart", lambda n: MultiAgentEnv(mode="bot_agent", bot_type="smart", num_envs=n, device=dev)),
]
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256,4096,16384").split(",")]
epws = (sys.argv[2] if len(sys.argv) > 2 else "auto,1,2,4,8,16,32").split(",")
for name, mk in cases:
    for n in sizes:
        for e in epws:
            if e == "auto":
                os.environ.pop("AT_SIM_EPW", None)
            else:
                os.environ["AT_SIM_EPW"] = e
            run(name, lambda: mk(n))
