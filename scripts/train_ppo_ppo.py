#!/usr/bin/env python
"""Batched counterpart of the reference's train_ppo_ppo.py (1v1 self-play: one PPO policy per tank):

    python scripts/train_ppo_ppo.py --num-envs 16384 --iterations 10

``MultiAgentEnv()`` (mode "agent", the "AI only" branch gaming_env.py:309-374) steps N battles at once; both tanks are
policy-driven (train_ppo_ppo.py:40-51: ``agents = [PPOAgentPPO(obs_dim, act_dim) for _ in range(num_tanks)]``), each
sees its own 388-column row through the trainer's per-tick normaliser (:91-93) and learns with GAE + the clipped
surrogate (:128-176).  The reference's early resets (``auto_reset_interval`` / negative-reward resets, :112-123) are not
mirrored: the batched env resets a battle when its episode ends.  Saves ``checkpoints/ppo_agent_<i>.pt`` = each agent's
``state_dict()`` (train_ppo_ppo.py:193-199), which the reference's inference.py loads."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--num-envs", type=int, default=16384)
    p.add_argument("--iterations", type=int, default=10)
    p.add_argument("--num-steps", type=int, default=32)
    p.add_argument("--num-epochs", type=int, default=4)
    p.add_argument("--num-minibatches", type=int, default=8)
    p.add_argument("--reward-mode", default="cumulative", choices=["cumulative", "delta"])
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints")
    a = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner
    torch.backends.cuda.matmul.allow_tf32 = True
    env = MultiAgentEnv(num_envs=a.num_envs, device=dev, seed=a.seed)
    env.render()
    num_tanks = env.num_tanks
    obs_dim = env.observation_space.shape[0] // num_tanks
    cfg = PPOConfig(num_steps=a.num_steps, num_epochs=a.num_epochs, num_minibatches=a.num_minibatches,
                    reward_mode=a.reward_mode)
    learner = PPOLearner(num_tanks, obs_dim, tuple(int(x) for x in env.action_space.nvec[:3]), cfg, dev, seed=a.seed)
    buf = RolloutBuffer(a.num_steps, a.num_envs, num_tanks, obs_dim, dev, keep_raw=False)
    runner = RolloutRunner(env, learner, buf)
    for it in range(a.iterations):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        runner.run(first=(it == 0))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        stats = learner.update(buf)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("iter %d  rollout %.1f M agent-steps/s (env + policies)  update %.2f s  reward/step %.3f  %s" % (
            it, a.num_envs * a.num_steps * num_tanks / (t1 - t0) / 1e6, t2 - t1, float(buf.rewards.mean()),
            "  ".join("agent %d pl %.3f vl %.3f ent %.3f" % (i, s["policy_loss"], s["value_loss"], s["entropy"])
                      for i, s in enumerate(stats))), flush=True)
    os.makedirs(a.ckpt_dir, exist_ok=True)
    for i, agent in enumerate(learner.agents):
        path = os.path.join(a.ckpt_dir, "ppo_agent_%d.pt" % i)
        torch.save(agent.state_dict(), path)
        print("[INFO] Saved model for Agent %d at %s" % (i, path))
    env.close()


if __name__ == "__main__":
    main()
