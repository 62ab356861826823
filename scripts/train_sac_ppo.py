#!/usr/bin/env python
"""Batched counterpart of the reference's train_sac_ppo.py (1v1: a PPO policy on tank 0 against a SAC policy on tank 1):

    python scripts/train_sac_ppo.py --num-envs 4096 --iterations 10

``MultiAgentEnv()`` (mode "agent") steps N battles at once.  As in the reference (train_sac_ppo.py:245-246, 283-330) the
PPO agent sees tank 0's raw 388-column row (no normaliser in this trainer), samples [rot, mov, shoot] and is updated
from an on-policy rollout with GAE + the clipped surrogate (:349-390); the SAC agent sees tank 1's row, its continuous
action is stored in the replay buffer (:322-328) and the twin critics / actor / temperature are updated from it
``sac_update_every`` times per iteration (:400-408).  The reference hands the SAC vector to ``env.step`` as is (its
``ContinuousToDiscreteWrapper`` line is commented out, :227-228), where non-integer components match no action code; here
it goes through ``ContinuousToDiscrete`` like in train_sac_sac.py so that the SAC tank acts.  The early resets
(:335-343: on done, every ``auto_reset_interval`` steps, or when a reward drops below ``-neg_reward_threshold``) become a
masked reset of exactly those envs, marked done for the PPO rollout like there.  Checkpoints: ``ppo_agent.pt`` and
``sac_agent.pt`` in the reference's layouts (:422-425).  No W&B / display manager."""
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
    p.add_argument("--ppo-num-steps", type=int, default=64, help="ticks per rollout (the reference: 512 single-env steps)")
    p.add_argument("--ppo-num-epochs", type=int, default=4)
    p.add_argument("--ppo-num-minibatches", type=int, default=8)
    p.add_argument("--sac-update-every", type=int, default=50)
    p.add_argument("--sac-batch-size", type=int, default=256)
    p.add_argument("--capacity", type=int, default=1_000_000)
    p.add_argument("--auto-reset-interval", type=int, default=10000)
    p.add_argument("--neg-reward-threshold", type=float, default=0.1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints")
    a = p.parse_args()

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(a.seed)
    torch.backends.cuda.matmul.allow_tf32 = True
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import (ContinuousToDiscrete, DeviceReplayBuffer, PPOConfig, PPOLearner, RolloutBuffer,
                                        SACAgent)
    env = MultiAgentEnv(num_envs=a.num_envs, device=dev, seed=a.seed)
    env.render()
    N, n, T = a.num_envs, env.num_tanks, a.ppo_num_steps
    if n < 2:
        raise ValueError("Environment must have at least 2 tanks for PPO vs SAC training.")     # train_sac_ppo.py:231
    obs_dim = env.observation_space.shape[0] // n
    nvec = tuple(int(v) for v in env.action_space.nvec[:3])
    sac_act_dim = int(env.action_space.shape[0])
    to_discrete = ContinuousToDiscrete(nvec)
    cfg = PPOConfig(num_steps=T, num_epochs=a.ppo_num_epochs, num_minibatches=a.ppo_num_minibatches, obs_norm="none")
    ppo = PPOLearner(1, obs_dim, nvec, cfg, dev, seed=a.seed)                       # tank 0
    sac = SACAgent(obs_dim, sac_act_dim, device=dev)                                # tank 1
    replay = DeviceReplayBuffer(a.capacity, obs_dim, sac_act_dim, dev)
    buf = RolloutBuffer(T, N, 1, obs_dim, dev, keep_raw=False)

    obs, _ = env.reset()
    obs = obs.view(N, n, obs_dim).clone()
    next_done = torch.zeros(N, dtype=torch.uint8, device=dev)
    actions = torch.zeros((N, n, 3), dtype=torch.int32, device=dev)
    global_step = 0
    for it in range(a.iterations):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        resets = 0
        for t in range(T):
            global_step += 1
            buf.obs[t].copy_(obs[:, 0:1])
            buf.dones[t].copy_(next_done)
            act, lp, val = ppo.act(buf.obs[t])
            buf.actions[t].copy_(act)
            buf.logprobs[t].copy_(lp)
            buf.values[t].copy_(val)
            cont = sac.select_action(obs[:, 1])
            actions[:, 0] = act[:, 0]
            actions[:, 1] = to_discrete(cont[..., :3])
            nxt, rew, done, _, _ = env.step(actions)
            nxt = nxt.view(N, n, obs_dim)
            buf.env_rewards[t].copy_(rew[:, 0:1])
            buf.rewards[t].copy_(rew[:, 0:1])
            replay.push(obs[:, 1], cont, rew[:, 1], nxt[:, 1], done.to(torch.float32))
            early = (rew < -a.neg_reward_threshold).any(dim=1) & (done == 0)
            if global_step % a.auto_reset_interval == 0:
                early = torch.ones_like(early)
            obs = nxt.clone()
            next_done = done.clone()
            if bool(early.any()):
                resets += int(early.sum())
                fresh = env.core.reset(early).view(N, n, obs_dim)
                obs[early] = fresh[early]
                next_done |= early.to(torch.uint8)                                  # "mark done", train_sac_ppo.py:343
        buf.obs[T].copy_(obs[:, 0:1])
        buf.dones[T].zero_()                                                        # train_sac_ppo.py:346
        buf.values[T].copy_(ppo.values(buf.obs[T]))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        (ps,) = ppo.update(buf)
        losses = []
        for _ in range(a.sac_update_every):
            if len(replay) >= a.sac_batch_size:
                losses.append(sac.update(replay, a.sac_batch_size))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        mean = [sum(l[k] for l in losses) / max(len(losses), 1) for k in range(4)]
        print("iter %d  %.2f M agent-steps/s collected  %d early resets  ppo pl %.3f vl %.3f ent %.3f  sac actor %.3f critic %.3f / %.3f "
              "alpha %.3f  updates %.2f s" % (it, N * n * T / (t1 - t0) / 1e6, resets, ps["policy_loss"], ps["value_loss"],
                                             ps["entropy"], *mean, t2 - t1), flush=True)
    os.makedirs(a.ckpt_dir, exist_ok=True)
    torch.save(ppo.agents[0].state_dict(), os.path.join(a.ckpt_dir, "ppo_agent.pt"))
    torch.save(sac.state_dict(), os.path.join(a.ckpt_dir, "sac_agent.pt"))
    print("[INFO] Saved models at %s" % a.ckpt_dir)
    env.close()


if __name__ == "__main__":
    main()
