#!/usr/bin/env python
"""Batched counterpart of the reference's train_ppo_cycle.py (one PPO agent against a rotation of
scripted bots under a weakness curriculum):

    python scripts/train_ppo_cycle.py --num-envs 4096 --updates 40 --rotation-strategy fixed

What is kept from the reference (train_ppo_cycle.py:24-92, 244-256, 258-445):

* ``CycleTrainingConfig``: the bot list, the rotation strategies "random" / "fixed" / "adaptive" (harder bots - lower
  rolling win rate - are drawn more often, :111-120), ``switch_every``, the rolling win-rate window, the win-rate
  threshold that gates the victory bonus (:341-350) and the "linear" / "sigmoid" weakness schedule (:82-92);
* the env is ``MultiAgentEnv(mode='bot_agent', type='train', bot_type=...)``, the agent is tank 1, the opponent is
  switched by ``set_bot_type`` (:305) and its weakness follows the schedule (``set_weakness`` -> ``at_set_weakness``,
  in place);
* GAE + clipped surrogate; observations go through ONE persistent ``RunningMeanStd`` updated with every batch of
  rows the agent sees (:279, :319-320; the fresh-normaliser-per-tick quirk belongs to train_ppo_bot.py /
  train_multi_ppo.py, not to this trainer), evaluation keeps its own (:102, :155); the checkpoint is the agent's
  ``state_dict()`` at ``checkpoints/single_ppo_cycle/ppo_agent_cycle.pt`` (:438-441), loadable by ``inference.py``;
* ``run_evaluation`` (:122-206): win rate and mean reward of the agent against every bot type.

What differs because N battles run at once: one env handle and rollout buffer per bot type (the running normaliser
keeps host-side state, so these rollouts are issued kernel by kernel, not as a CUDA graph); an episode's outcome is ``info["winner"] == 1`` at its terminal tick exactly as in the
reference (:341, :357) - the step latches the finished episode's info before its auto-reset (``at_get_terminal_info``),
and the rollout stores it per tick (``RolloutBuffer(keep_winner=True)``).  W&B logging and the video recorder are out
of scope."""
import argparse
import collections
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROTATION_STRATEGIES = ["random", "fixed", "adaptive"]          # train_ppo_cycle.py:21


class CycleTrainingConfig:
    """train_ppo_cycle.py:24-92 (same names and defaults; ``num_steps`` is ticks per rollout of N envs)."""

    def __init__(self, bot_types=None, rotation_strategy="random", switch_every=10000, rolling_window_size=100,
                 eval_frequency=1000, eval_episodes=50, win_rate_threshold=0.8, reward_threshold_percentage=0.3,
                 learning_rate=3e-4, num_steps=64, num_epochs=4, total_timesteps=100000, gamma=0.99, gae_lambda=0.95,
                 clip_coef=0.1, ent_coef=0.02, vf_coef=0.3, max_grad_norm=0.3, weakness_start=0.1, weakness_end=0.8,
                 weakness_schedule="linear", weakness_sigmoid_steepness=10.0):
        self.bot_types = list(bot_types or ["random", "aggressive", "defensive", "dodge"])
        if rotation_strategy not in ROTATION_STRATEGIES:
            raise ValueError("rotation_strategy must be one of %s" % ROTATION_STRATEGIES)
        if weakness_schedule not in ["linear", "sigmoid"]:
            raise ValueError("weakness_schedule must be either 'linear' or 'sigmoid'")
        self.rotation_strategy, self.switch_every = rotation_strategy, switch_every
        self.rolling_window_size, self.eval_frequency, self.eval_episodes = rolling_window_size, eval_frequency, eval_episodes
        self.win_rate_threshold, self.reward_threshold_percentage = win_rate_threshold, reward_threshold_percentage
        self.learning_rate, self.num_steps, self.num_epochs, self.total_timesteps = learning_rate, num_steps, num_epochs, total_timesteps
        self.gamma, self.gae_lambda, self.clip_coef, self.ent_coef = gamma, gae_lambda, clip_coef, ent_coef
        self.vf_coef, self.max_grad_norm = vf_coef, max_grad_norm
        self.weakness_start, self.weakness_end = weakness_start, weakness_end
        self.weakness_schedule, self.weakness_sigmoid_steepness = weakness_schedule, weakness_sigmoid_steepness

    def get_current_weakness(self, current_step):
        progress = min(1.0, current_step / self.total_timesteps)
        if self.weakness_schedule == "linear":
            return self.weakness_start + (self.weakness_end - self.weakness_start) * progress
        x = (progress - 0.5) * self.weakness_sigmoid_steepness
        return self.weakness_start + (self.weakness_end - self.weakness_start) / (1.0 + math.exp(-x))


class WinTracker:
    """Rolling win rates per bot type (train_ppo_cycle.py:97-120, 275-290).  One entry per finished batch of episodes:
    the reference appends one 0/1 per episode; here a rollout contributes its wins / episodes ratio once."""

    def __init__(self, config):
        self.config = config
        self.rates = {b: collections.deque(maxlen=config.rolling_window_size) for b in config.bot_types}

    def update(self, bot_type, wins, episodes):
        if episodes > 0:
            self.rates[bot_type].append(wins / episodes)

    def get_win_rate(self, bot_type):
        r = self.rates[bot_type]
        return sum(r) / len(r) if r else 0.0

    def get_bot_probabilities(self):
        bots = self.config.bot_types
        if not all(len(self.rates[b]) > 0 for b in bots):
            return {b: 1.0 / len(bots) for b in bots}
        inv = {b: 1.0 - self.get_win_rate(b) for b in bots}              # harder bots more often (:116-120)
        total = sum(inv.values())
        if total <= 0:
            return {b: 1.0 / len(bots) for b in bots}
        return {b: v / total for b, v in inv.items()}


def select_next_bot(config, tracker, current_bot, rng):
    """train_ppo_cycle.py:244-256."""
    if config.rotation_strategy == "random":
        return str(rng.choice(config.bot_types))
    if config.rotation_strategy == "fixed":
        if current_bot is None:
            return config.bot_types[0]
        return config.bot_types[(config.bot_types.index(current_bot) + 1) % len(config.bot_types)]
    probs = tracker.get_bot_probabilities()
    return str(rng.choice(config.bot_types, p=[probs[b] for b in config.bot_types]))


def rollout_outcomes(buf):
    """(wins, episodes, won[T, N]) of the episodes that ended inside a filled rollout buffer: the agent (tank 1) won
    where the terminal tick's ``info["winner"]`` is 1 (train_ppo_cycle.py:341, 357)."""
    ended = buf.dones[1:].bool()
    won = ended & (buf.winner == 1)
    return int(won.sum()), int(ended.sum()), won


@torch.no_grad()
def run_evaluation(learner, bot_types, num_envs, ticks, device, seed=12345, weakness=1.0, eval_norm=None):
    """train_ppo_cycle.py:122-206: the agent against every bot type; win = ``info["winner"] == 1`` at the terminal
    tick (:181).  ``num_envs`` battles for ``ticks`` ticks each instead of ``eval_episodes`` sequential episodes; the
    finished battles are reset by the caller after reading info, the reference's own order.  Observations go through
    the evaluator's OWN running normaliser (:102, :155), not the trainer's: pass one ``RunningMeanStd`` as
    ``eval_norm`` to keep it across evaluations like the reference's evaluator object does."""
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner.ppo import SingleAgentView
    from alphatank_b200.learner.ppo_utils import RunningMeanStd
    results = {}
    for bot in bot_types:
        env = MultiAgentEnv(mode="bot_agent", type="train", bot_type=bot, weakness=weakness, num_envs=num_envs,
                            device=device, seed=seed, auto_reset=False)
        view = SingleAgentView(env).core
        D = view.D
        raw = torch.zeros((num_envs, D), dtype=torch.float32, device=device)
        rew = torch.zeros((num_envs,), dtype=torch.float32, device=device)
        done = torch.zeros((num_envs,), dtype=torch.uint8, device=device)
        view.reset(obs_out=raw)
        if eval_norm is None:
            eval_norm = RunningMeanStd((D,), device=device)
        wins = episodes = 0
        reward_sum = 0.0
        for _ in range(ticks):
            eval_norm.update(raw)
            acts, _, _ = learner.act(eval_norm.normalize(raw).view(num_envs, 1, D))
            view.step(acts, obs_out=raw, rewards_out=rew, done_out=done)
            d = done.bool()
            if bool(d.any()):
                w = env.core.get_info()["winner"]
                wins += int((w[d] == 1).sum())
                episodes += int(d.sum())
                reward_sum += float(rew[d].sum())
                view.reset(mask=done, obs_out=raw)
        results[bot] = {"win_rate": wins / max(episodes, 1), "avg_reward": reward_sum / max(episodes, 1),
                        "episodes": episodes}
        env.close()
    return results


def train_cycle(config, num_envs=4096, updates=None, seed=0, eval_envs=1024, eval_ticks=400,
                ckpt_dir="checkpoints/single_ppo_cycle", device="cuda:0", log=print):
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner
    from alphatank_b200.learner.ppo import SingleAgentView
    dev = torch.device(device)
    torch.backends.cuda.matmul.allow_tf32 = True
    rng = np.random.default_rng(seed)
    T = config.num_steps
    num_updates = updates if updates is not None else max(1, config.total_timesteps // (T * num_envs))
    switch_updates = max(1, config.switch_every // (T * num_envs))
    eval_updates = max(1, config.eval_frequency // (T * num_envs)) if config.eval_frequency else 0
    cfg = PPOConfig(learning_rate=config.learning_rate, gamma=config.gamma, gae_lambda=config.gae_lambda,
                    clip_coef=config.clip_coef, ent_coef=config.ent_coef, vf_coef=config.vf_coef,
                    max_grad_norm=config.max_grad_norm, num_steps=T, num_epochs=config.num_epochs,
                    obs_norm="running")       # train_ppo_cycle.py:279, 319-320: one persistent RunningMeanStd
    learner = tracker = eval_norm = None
    slots = {}                                                           # bot type -> env, buffer, runner, state

    def slot(bot):
        nonlocal learner
        if bot not in slots:
            env = MultiAgentEnv(mode="bot_agent", type="train", bot_type=bot, weakness=config.weakness_start,
                                num_envs=num_envs, device=dev, seed=seed + 1000 * len(slots))
            env.render()
            obs_dim = env.observation_space.shape[0] // env.num_tanks
            if learner is None:
                learner = PPOLearner(1, obs_dim, tuple(int(x) for x in env.action_space.nvec[:3]), cfg, dev, seed=seed)
            buf = RolloutBuffer(T, num_envs, 1, obs_dim, dev, keep_raw=True, keep_winner=True)   # (the running normaliser reads the raw rows)
            slots[bot] = {"env": env, "buf": buf, "runner": RolloutRunner(SingleAgentView(env), learner, buf),
                          "first": True}
        return slots[bot]

    current_bot = config.bot_types[0]
    slot(current_bot)
    tracker = WinTracker(config)
    history = []
    for update in range(1, num_updates + 1):
        current_step = update * T * num_envs
        weakness = config.get_current_weakness(current_step)
        if update % switch_updates == 0:
            current_bot = select_next_bot(config, tracker, current_bot, rng)
        s = slot(current_bot)
        s["env"].set_weakness(weakness)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        s["runner"].run(first=s["first"])
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        buf = s["buf"]
        wins, episodes, won = rollout_outcomes(buf)
        s["first"] = False
        # victory bonus, gated by the curriculum (train_ppo_cycle.py:341-350)
        rates = [tracker.get_win_rate(b) for b in config.bot_types]
        if update < num_updates * config.reward_threshold_percentage:
            weight = sum(r >= config.win_rate_threshold for r in rates) / len(config.bot_types)
        else:
            weight = 1.0 if all(r >= config.win_rate_threshold for r in rates) else 0.0
        if weight > 0:
            buf.rewards.add_(won.unsqueeze(-1).to(torch.float32) * weight)
        tracker.update(current_bot, wins, episodes)
        (stats,) = learner.update(buf)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        rec = {"update": update, "bot": current_bot, "weakness": weakness, "episodes": episodes,
               "win_rate": tracker.get_win_rate(current_bot), "agent_steps_per_s": num_envs * T / (t1 - t0)}
        history.append(rec)
        log("update %d  bot %-10s weakness %.2f  %6d episodes  win rate %.2f  rollout %.1f M agent-steps/s  update %.2f s  "
            "pl %.3f vl %.3f ent %.3f" % (update, current_bot, weakness, episodes, rec["win_rate"],
                                          rec["agent_steps_per_s"] / 1e6, t2 - t1, stats["policy_loss"],
                                          stats["value_loss"], stats["entropy"]))
        if eval_updates and update % eval_updates == 0:
            if eval_norm is None:
                from alphatank_b200.learner.ppo_utils import RunningMeanStd
                eval_norm = RunningMeanStd((learner.D,), device=dev)
            ev = run_evaluation(learner, config.bot_types, eval_envs, eval_ticks, dev, eval_norm=eval_norm)
            rec["eval"] = ev
            log("  eval: " + "  ".join("%s %.2f (%d ep)" % (b, r["win_rate"], r["episodes"]) for b, r in ev.items()))
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, "ppo_agent_cycle.pt")
    torch.save(learner.agents[0].state_dict(), path)
    log("[INFO] Saved model at %s" % path)
    for s in slots.values():
        s["env"].close()
    return learner, history, path


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--bot-types", nargs="+", default=["random", "aggressive", "defensive", "dodge"])
    p.add_argument("--rotation-strategy", default="fixed", choices=ROTATION_STRATEGIES)
    p.add_argument("--num-envs", type=int, default=4096)
    p.add_argument("--num-steps", type=int, default=64)
    p.add_argument("--updates", type=int, default=20)
    p.add_argument("--switch-updates", type=int, default=2, help="rollouts per bot before the rotation moves on")
    p.add_argument("--eval-updates", type=int, default=10)
    p.add_argument("--eval-envs", type=int, default=1024)
    p.add_argument("--eval-ticks", type=int, default=400)
    p.add_argument("--weakness-start", type=float, default=0.1)
    p.add_argument("--weakness-end", type=float, default=0.8)
    p.add_argument("--weakness-schedule", default="linear", choices=["linear", "sigmoid"])
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ckpt-dir", default="checkpoints/single_ppo_cycle")
    a = p.parse_args()
    torch.cuda.set_device(0)
    per_update = a.num_steps * a.num_envs
    config = CycleTrainingConfig(bot_types=a.bot_types, rotation_strategy=a.rotation_strategy,
                                 switch_every=a.switch_updates * per_update, eval_frequency=a.eval_updates * per_update,
                                 num_steps=a.num_steps, total_timesteps=a.updates * per_update,
                                 weakness_start=a.weakness_start, weakness_end=a.weakness_end,
                                 weakness_schedule=a.weakness_schedule)
    train_cycle(config, num_envs=a.num_envs, updates=a.updates, seed=a.seed, eval_envs=a.eval_envs,
                eval_ticks=a.eval_ticks, ckpt_dir=a.ckpt_dir)


if __name__ == "__main__":
    main()
