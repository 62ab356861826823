#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations other than the bench line (they are parity-test cases; this is for
the DESIGN / profiles tables).  Device-resident random actions, 400 burn-in ticks, CUDA events.  Development tooling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv  # noqa: E402
from alphatank_b200.configs.config_teams import team_vs_bot_configs  # noqa: E402

dev = torch.device("cuda:0")


def run(name, env, agents, steps=200, burn=400):
    core = env.core
    env.reset()
    A = core.A
    acts = None
    if A:
        acts = torch.empty((64, core.num_envs, A, 3), dtype=torch.int32, device=dev)
        for t in range(64):
            core.synthetic_actions(t, acts[t])
    for t in range(burn):
        core.step(acts[t % 64] if A else None)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for t in range(steps):
        core.step(acts[t % 64] if A else None)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    n = core.num_envs
    print("%-58s %8d envs  %.4f ms/tick  %7.1f M env-steps/s  %7.1f M agent-steps/s  obs %d x %d" % (
        name, n, ms, n / ms / 1e3, n * agents / ms / 1e3, core.R, core.D), flush=True)
    env.close() if hasattr(env, "close") else None


run("2v2 team_vs_bot, 256 envs (smoke size)", MultiAgentTeamEnv(team_vs_bot_configs, num_envs=256, device=dev), 2)
run("2v2 team_vs_bot, 4096 envs", MultiAgentTeamEnv(team_vs_bot_configs, num_envs=4096, device=dev), 2)
run("config 2: 1v1 bot_agent vs smart bot", MultiAgentEnv(mode="bot_agent", bot_type="smart", num_envs=4096, device=dev), 1)
run("config 2 at 65 536 envs", MultiAgentEnv(mode="bot_agent", bot_type="smart", num_envs=65536, device=dev), 1)
run("config 1 batched: 1v1 agent mode", MultiAgentEnv(mode="agent", num_envs=65536, device=dev), 2)
run("config 3: 2v2 team_vs_bot (the bench line)", MultiAgentTeamEnv(team_vs_bot_configs, num_envs=65536, device=dev), 2)
run("config 4: arena defensive vs dodge, TERMINATE_TIME 4096",
    MultiAgentEnv(mode="bot_vs_bot", bot_types=("defensive", "dodge"), num_envs=262144, device=dev, terminate_time=4096), 2)
run("config 5 per GPU: 2v2 team_vs_bot", MultiAgentTeamEnv(team_vs_bot_configs, num_envs=131072, device=dev), 2)
