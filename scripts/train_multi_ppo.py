#!/usr/bin/env python
"""Batched counterpart of the reference's train_multi_ppo.py (same CLI names where they exist):

    python scripts/train_multi_ppo.py --experiment_name 2a_vs_2b --config_var team_vs_bot_configs --num-envs 65536
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_multi_ppo.py ...

One process per GPU; rank r simulates env ids [r*E, (r+1)*E) (no collective on the step path) and the A policies are
data parallel: one NCCL all-reduce of the flattened gradient per optimiser step.  Writes
``checkpoints/team_ppo/<experiment_name>.pth`` in the reference's layout ({'team_config', '<TankName>': state_dict}),
which the reference's inference_multi.py loads unchanged.  No W&B / video recorder (out of scope)."""
import argparse
import importlib
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--experiment_name", type=str, default="2a_vs_2b")
    p.add_argument("--config_var", type=str, default="team_vs_bot_configs")
    p.add_argument("--num-envs", type=int, default=16384, help="envs per GPU")
    p.add_argument("--iterations", type=int, default=10)
    p.add_argument("--num-steps", type=int, default=32)
    p.add_argument("--num-epochs", type=int, default=4)
    p.add_argument("--num-minibatches", type=int, default=8)
    p.add_argument("--reward-mode", default="cumulative", choices=["cumulative", "delta"])
    p.add_argument("--obs-norm", default="tick", choices=["tick", "running", "none"])
    p.add_argument("--no-graph", action="store_true", help="issue the rollout kernel by kernel instead of one CUDA graph")
    p.add_argument("--obs-fp16", action="store_true",
                   help="keep the tick-normalised rows as fp16 (at_step_norm16 + the fp16 tensor-core policy kernel; needs --obs-norm tick)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints/team_ppo")
    a = p.parse_args()

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from alphatank_b200 import MultiAgentTeamEnv
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner, save_team_checkpoint
    torch.backends.cuda.matmul.allow_tf32 = True      # the policies' GEMMs on the tensor cores (cuBLAS TF32)
    cfgs = importlib.import_module("alphatank_b200.configs.config_teams")
    if not hasattr(cfgs, a.config_var):
        raise ValueError("Variable %s not found in configs/config_teams.py" % a.config_var)     # train_multi_ppo.py:226
    game_configs = getattr(cfgs, a.config_var)

    env = MultiAgentTeamEnv(game_configs, num_envs=a.num_envs, device=dev, seed=a.seed, env_id_base=rank * a.num_envs)
    env.render()
    names = env.get_observation_order()
    A = len(names)
    D = env.observation_space.shape[0] // A
    cfg = PPOConfig(num_steps=a.num_steps, num_epochs=a.num_epochs, num_minibatches=a.num_minibatches,
                    reward_mode=a.reward_mode, obs_norm=a.obs_norm)
    learner = PPOLearner(A, D, tuple(int(x) for x in env.action_space.nvec[:3]), cfg, dev, seed=a.seed)
    buf = RolloutBuffer(a.num_steps, a.num_envs, A, D, dev, keep_raw=a.obs_norm != "tick",
                        obs_dtype=torch.float16 if (a.obs_fp16 and a.obs_norm == "tick") else torch.float32)
    runner = RolloutRunner(env, learner, buf, use_graph=not a.no_graph)
    for it in range(a.iterations):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        runner.run(first=(it == 0))
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        stats = learner.update(buf)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        if rank == 0:
            sps = a.num_envs * world * a.num_steps * A / (t1 - t0)
            print("iter %d  rollout %.1f M agent-steps/s (env + policy)  update %.2f s  reward/step %.3f  %s" % (
                it, sps / 1e6, t2 - t1, float(buf.rewards.mean()),
                "  ".join("%s pl %.3f vl %.3f ent %.3f" % (n, s["policy_loss"], s["value_loss"], s["entropy"])
                          for n, s in zip(names, stats))), flush=True)
    if rank == 0:
        path = save_team_checkpoint(os.path.join(a.ckpt_dir, a.experiment_name + ".pth"), game_configs, names, learner.agents)
        print("[INFO] Saved model and config at %s" % path)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
