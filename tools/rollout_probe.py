#!/usr/bin/env python
"""Times the PPO rollout (policies in the loop) at bench size and, with --list, prints the per-kernel device time of
one rollout from torch.profiler.  Development tooling (profiles/*.md quote its output)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--num-envs", type=int, default=65536)
    p.add_argument("--num-steps", type=int, default=16)
    p.add_argument("--iters", type=int, default=6)
    p.add_argument("--list", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--half", action="store_true", help="fp16 observations (at_step_norm16 + at_policy_forward16)")
    a = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from alphatank_b200 import MultiAgentTeamEnv
    from alphatank_b200.configs.config_teams import team_vs_bot_configs
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner
    torch.backends.cuda.matmul.allow_tf32 = True
    env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=a.num_envs, device=dev, seed=0)
    names = env.get_observation_order()
    A = len(names)
    D = env.observation_space.shape[0] // A
    learner = PPOLearner(A, D, (3, 3, 2), PPOConfig(num_steps=a.num_steps), dev, seed=0)
    buf = RolloutBuffer(a.num_steps, a.num_envs, A, D, dev, keep_raw=False, obs_dtype=torch.float16 if a.half else torch.float32)
    runner = RolloutRunner(env, learner, buf, use_graph=not a.no_graph)
    runner.run(first=True)
    for _ in range(12):          # into the steady-state mix of episode phases
        runner.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        runner.run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.iters
    print("rollout (%s observations): %.3f ms per tick, %.1f M agent-steps/s (env + policies, T=%d, %d envs)" % (
        "fp16" if a.half else "fp32", 1e3 * dt / a.num_steps, a.num_envs * a.num_steps * A / dt / 1e6, a.num_steps, a.num_envs), flush=True)
    if a.list:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            runner.run()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        print("device time of one rollout: %.3f ms (%.3f ms per tick)" % (tot / 1e3, tot / 1e3 / a.num_steps))
        for e in rows[:30]:
            print("%5.1f%%  n=%4d  avg %8.1f us  %s" % (100 * e.device_time_total / tot, e.count,
                                                       e.device_time_total / max(e.count, 1), e.key[:110]))


if __name__ == "__main__":
    main()
