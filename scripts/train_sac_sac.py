#!/usr/bin/env python
"""Batched counterpart of the reference's train_sac_sac.py (SAC vs SAC, 1v1 agent mode):

    python scripts/train_sac_sac.py --num-envs 4096 --iterations 20

N battles step together on the GPU; everything the reference does per transition is done per batch of N transitions on
the device: the two policies act on ``[N, 388]`` observation blocks (``SACAgent.select_action``), the continuous
actions go through ``ContinuousToDiscrete`` (the wrapper's rounding, models/sac_utils.py:22-33 - only the first three of
an agent's six components reach its tank, like there), the transitions go into ``DeviceReplayBuffer`` ring tensors in
HBM, and the twin-critic update (models/sac_utils.py:160-220) samples from them.  The reference's early resets
(train_sac_sac.py:102-108: on done, every ``auto_reset_interval`` steps, or when a reward drops below
``-neg_reward_threshold``) become a masked ``reset`` of exactly those envs.  Checkpoints: ``sac_agent_<i>.pt`` holding
``SACAgent.state_dict()`` in the reference's layout (train_sac_sac.py:145-150).  No W&B / video recorder."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--num-envs", type=int, default=4096)
    p.add_argument("--iterations", type=int, default=10)
    p.add_argument("--num-steps", type=int, default=64, help="ticks per iteration (the reference: 512 single-env steps)")
    p.add_argument("--start-steps", type=int, default=16, help="ticks of uniform random actions before the policies act")
    p.add_argument("--update-every", type=int, default=50, help="SAC updates per agent and iteration")
    p.add_argument("--batch-size", type=int, default=256)
    p.add_argument("--capacity", type=int, default=1_000_000)
    p.add_argument("--auto-reset-interval", type=int, default=20000)
    p.add_argument("--neg-reward-threshold", type=float, default=0.1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints/single_sac_vs_sac")
    a = p.parse_args()

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(a.seed)
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import ContinuousToDiscrete, DeviceReplayBuffer, SACAgent
    env = MultiAgentEnv(num_envs=a.num_envs, device=dev, seed=a.seed, auto_reset=True)      # mode "agent": both tanks learn
    env.render()
    N, n = a.num_envs, env.num_tanks
    obs_dim = env.observation_space.shape[0] // n
    action_dim = int(env.action_space.shape[0])                                             # 6: [3, 3, 2] per tank
    to_discrete = ContinuousToDiscrete(tuple(int(v) for v in env.action_space.nvec[:3]))
    agents = [SACAgent(obs_dim, action_dim, device=dev) for _ in range(n)]
    buffers = [DeviceReplayBuffer(a.capacity, obs_dim, action_dim, dev) for _ in range(n)]

    obs, _ = env.reset()
    obs = obs.view(N, n, obs_dim).clone()
    global_step = 0
    for it in range(a.iterations):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        resets = 0
        for _ in range(a.num_steps):
            global_step += 1
            if global_step <= a.start_steps:
                cont = torch.rand((N, n, action_dim), device=dev)
            else:
                cont = torch.stack([ag.select_action(obs[:, i]) for i, ag in enumerate(agents)], dim=1)
            nxt, rew, done, _, _ = env.step(to_discrete(cont[..., :3]))
            nxt = nxt.view(N, n, obs_dim)
            d = done.to(torch.float32)
            for i in range(n):
                buffers[i].push(obs[:, i], cont[:, i], rew[:, i], nxt[:, i], d)
            # finished envs were reset in place by the step (auto_reset); the reference also resets early
            early = (rew < -a.neg_reward_threshold).any(dim=1) & (done == 0)
            if global_step % a.auto_reset_interval == 0:
                early = torch.ones_like(early)
            obs = nxt.clone()
            if bool(early.any()):
                resets += int(early.sum())
                fresh = env.core.reset(early).view(N, n, obs_dim)
                obs[early] = fresh[early]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        losses = []
        for _ in range(a.update_every):
            for i, ag in enumerate(agents):
                if len(buffers[i]) >= a.batch_size:
                    losses.append(ag.update(buffers[i], a.batch_size))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        mean = [sum(l[k] for l in losses) / max(len(losses), 1) for k in range(4)]
        print("iter %d  %.2f M agent-steps/s collected  %d early resets  %d updates in %.2f s  actor %.3f critic %.3f / %.3f alpha %.3f"
              % (it, N * n * a.num_steps / (t1 - t0) / 1e6, resets, len(losses), t2 - t1, *mean), flush=True)
    os.makedirs(a.ckpt_dir, exist_ok=True)
    for i, ag in enumerate(agents):
        path = os.path.join(a.ckpt_dir, "sac_agent_%d.pt" % i)
        torch.save(ag.state_dict(), path)
        print("[INFO] Saved model for Agent %d at %s" % (i, path))
    env.close()


if __name__ == "__main__":
    main()
