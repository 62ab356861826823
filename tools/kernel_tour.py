#!/usr/bin/env python
"""Small tour of every kernel the library launches: the step in its plain / normalised / host-buffer forms on
four env kinds (partial last block, per-env mazes), masked resets, the info kernels, state snapshots, and the
learner-side kernels - with torch's caching allocator off, so every buffer is its own cudaMalloc.  (Written for
compute-sanitizer; that tool is closed on this GPU pool, so the tour only checks that every launch succeeds.)
Development tooling."""
import os
import sys

os.environ.setdefault("PYTORCH_NO_CUDA_MEMORY_CACHING", "1")   # one cudaMalloc per tensor: out-of-bounds stores are visible
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv  # noqa: E402
from alphatank_b200.configs.config_teams import team_vs_bot_configs, team_vs_bot_harder_configs  # noqa: E402

dev = torch.device("cuda:0")
TICKS = int(os.environ.get("TOUR_TICKS", "24"))


def tour(name, env, norm_ok=True):
    core = env.core
    N, A = core.num_envs, core.A
    env.reset()
    acts = torch.zeros((N, max(A, 1), 3), dtype=torch.int32, device=dev)
    norm = torch.empty((N, core.R * core.D), dtype=torch.float32, device=dev)
    done_total = 0
    for t in range(TICKS):
        if A:
            core.synthetic_actions(t, acts)
        a = acts if A else None
        if t % 3 == 0:
            _, _, d = core.step(a)
        elif t % 3 == 1 and core.R > 0 and norm_ok:
            _, _, d = core.step_norm(a, norm)
        else:
            _, _, d = core.step_host(a.cpu().pin_memory() if A else None)
            d = torch.as_tensor(d)
        done_total += int(d.sum())
        if t % 5 == 0:
            core.get_info(terminal=True)
            core.get_info()
        if t % 7 == 0:
            m = torch.zeros(N, dtype=torch.uint8, device=dev)
            m[::3] = 1
            core.reset(m)
    core.get_state(0, min(N, 5))
    torch.cuda.synchronize()
    print("%-28s %5d envs  %d ticks  %d episodes ended" % (name, N, TICKS, done_total), flush=True)
    env.close()


tour("team 2v2 (spec 3)", MultiAgentTeamEnv(team_vs_bot_configs, num_envs=200, device=dev, seed=1))
tour("team harder, battlefield", MultiAgentTeamEnv(team_vs_bot_harder_configs, num_envs=70, device=dev, seed=2, map="battlefield"))
tour("1v1 bot_agent smart, maze", MultiAgentEnv(mode="bot_agent", bot_type="smart", num_envs=100, device=dev, seed=3, map="maze"))
tour("arena defensive vs dodge", MultiAgentEnv(mode="bot_vs_bot", bot_types=("defensive", "dodge"), num_envs=65, device=dev, seed=4,
                                               terminate_time=16))

# learner-side kernels: tick normaliser, sampling, policy tail / mid kernels through one small rollout
from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, collect_rollout  # noqa: E402
env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=256, device=dev, seed=6)
learner = PPOLearner(2, 528, (3, 3, 2), PPOConfig(num_steps=4), dev, seed=0)
buf = RolloutBuffer(4, 256, 2, 528, dev, keep_raw=False, keep_winner=True)
collect_rollout(env, learner, buf, first=True)
collect_rollout(env, learner, buf, first=False)
from alphatank_b200.learner.ppo_utils import tick_normalize  # noqa: E402
tick_normalize(torch.rand((300, 2, 528), device=dev))
torch.cuda.synchronize()
print("learner kernels ok", flush=True)
env.close()
