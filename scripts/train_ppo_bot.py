#!/usr/bin/env python
"""Batched counterpart of the reference's train_ppo_bot.py (BASELINE config 2: one PPO agent against a scripted bot):

    python scripts/train_ppo_bot.py --bot-type smart --num-envs 4096

``MultiAgentEnv(mode='bot_agent', type='train', bot_type=...)`` steps N battles at once with the bot driven on the device;
the agent is tank 1 (train_ppo_bot.py:35), sees its own 388-column row normalised by the trainer's per-tick
``RunningMeanStd`` (train_ppo_bot.py:97-99: with a single row that is x / (1 + eps) scaled by 1 / sqrt(1 + x^2 eps /
(1 + eps)^2)), and learns with the reference's GAE + clipped surrogate.  The reference's early resets (on a negative
reward, train_ppo_bot.py:121-129) are not mirrored here: the batched env resets an env when its episode ends.  Saves
``checkpoints/single_ppo_vs_bots/ppo_agent_vs_<bot>.pt`` = the agent's ``state_dict()`` (train_ppo_bot.py:196-200), which
the reference's inference.py loads."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--bot-type", default="smart", choices=["smart", "random", "aggressive", "defensive", "dodge"])
    p.add_argument("--num-envs", type=int, default=4096)
    p.add_argument("--iterations", type=int, default=10)
    p.add_argument("--num-steps", type=int, default=64)
    p.add_argument("--num-epochs", type=int, default=4)
    p.add_argument("--num-minibatches", type=int, default=8)
    p.add_argument("--weakness", type=float, default=1.0)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints/single_ppo_vs_bots")
    a = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner
    from alphatank_b200.learner.ppo import SingleAgentView
    torch.backends.cuda.matmul.allow_tf32 = True
    env = MultiAgentEnv(mode="bot_agent", type="train", bot_type=a.bot_type, weakness=a.weakness, num_envs=a.num_envs,
                        device=dev, seed=a.seed)
    env.render()
    obs_dim = env.observation_space.shape[0] // env.num_tanks
    cfg = PPOConfig(num_steps=a.num_steps, num_epochs=a.num_epochs, num_minibatches=a.num_minibatches)
    learner = PPOLearner(1, obs_dim, tuple(int(x) for x in env.action_space.nvec[:3]), cfg, dev, seed=a.seed)
    buf = RolloutBuffer(a.num_steps, a.num_envs, 1, obs_dim, dev, keep_raw=False)
    runner = RolloutRunner(SingleAgentView(env), learner, buf)
    for it in range(a.iterations):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        runner.run(first=(it == 0))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        (s,) = learner.update(buf)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("iter %d  rollout %.1f M agent-steps/s (env + bot + policy)  update %.2f s  reward/step %.3f  pl %.3f vl %.3f ent %.3f"
              % (it, a.num_envs * a.num_steps / (t1 - t0) / 1e6, t2 - t1, float(buf.rewards.mean()), s["policy_loss"],
                 s["value_loss"], s["entropy"]), flush=True)
    os.makedirs(a.ckpt_dir, exist_ok=True)
    path = os.path.join(a.ckpt_dir, "ppo_agent_vs_%s.pt" % a.bot_type)
    torch.save(learner.agents[0].state_dict(), path)
    print("[INFO] Saved model at %s" % path)
    env.close()


if __name__ == "__main__":
    main()
