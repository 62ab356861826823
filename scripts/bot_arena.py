#!/usr/bin/env python
"""Batched counterpart of the reference's bot_arena.py (BASELINE config 4: scripted bot against scripted bot):

    python scripts/bot_arena.py --bot1 defensive --bot2 dodge --num-envs 65536 --ticks 2000
    python scripts/bot_arena.py --all --num-envs 16384 --ticks 1500          # round robin of every pair

The reference opens a window, lets ``BotFactory.create_bot(bot1)`` drive tank 0 and ``create_bot(bot2)`` tank 1 of a
``GamingENV(mode="bot_vs_bot")``, both deciding on the pre-tick state (bot_arena.py:51-58), and resets as soon as one
of the tanks dies (bot_arena.py:61-70).  Here N such matches run at once in ``MultiAgentEnv(mode="bot_vs_bot",
bot_types=(bot1, bot2))`` with both controllers on the device and finished matches restarted inside the step; the
winner is ``info["winner"]`` of the terminal tick (first alive tank, gym_env.py:97; -1 = both died in the same tick),
which the step latches before its auto-reset (``at_get_terminal_info``).
``--terminate-time T`` gives the stress variant of BASELINE config 4 (immortal tanks, bullets saturate)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BOTS = ["smart", "random", "aggressive", "defensive", "dodge", "stationary"]   # BotFactory.BOT_TYPES (bot_factory.py:7-14)


def run_bot_matches(bot1_type, bot2_type, num_envs=4096, ticks=1000, device="cuda:0", seed=0, terminate_time="config",
                    map=None):
    """Runs ``num_envs`` matches side by side for ``ticks`` ticks; returns a dict of counters (python ints / floats)."""
    from alphatank_b200 import MultiAgentEnv
    dev = torch.device(device)
    env = MultiAgentEnv(mode="bot_vs_bot", bot_types=(bot1_type, bot2_type), num_envs=num_envs, device=dev, seed=seed,
                        terminate_time=terminate_time, map=map)
    env.reset()
    wins = torch.zeros(3, dtype=torch.int64, device=dev)              # bot1, bot2, nobody
    length = torch.zeros(num_envs, dtype=torch.int64, device=dev)
    total_len = torch.zeros((), dtype=torch.int64, device=dev)
    hits = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(ticks):
        _, _, done, _, _ = env.step()
        length += 1
        d = done.bool()
        gi = env.core.get_info(terminal=True)                         # = info of this step: the finished match's where done
        w = gi["winner"].long()
        idx = torch.where(w < 0, torch.full_like(w, 2), w)
        wins += torch.bincount(idx[d], minlength=3)[:3]
        total_len += length[d].sum()
        hits += gi["hits"].long()[d].sum(0)
        length.masked_fill_(d, 0)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    w = wins.tolist()
    episodes = sum(w)
    out = {"bot1": bot1_type, "bot2": bot2_type, "num_envs": num_envs, "ticks": ticks, "episodes": episodes,
           "bot1_wins": w[0], "bot2_wins": w[1], "draws": w[2],
           "mean_episode_ticks": float(total_len) / max(episodes, 1),
           "bot1_hits": int(hits[0]), "bot2_hits": int(hits[1]),
           "env_steps_per_s": num_envs * ticks / dt}
    env.close()
    return out


def _line(r):
    e = max(r["episodes"], 1)
    return "%-10s vs %-10s  %8d matches  %5.1f %% / %5.1f %% / %4.1f %% draws  %6.1f ticks per match  %6.1f M env-steps/s" % (
        r["bot1"], r["bot2"], r["episodes"], 100.0 * r["bot1_wins"] / e, 100.0 * r["bot2_wins"] / e,
        100.0 * r["draws"] / e, r["mean_episode_ticks"], r["env_steps_per_s"] / 1e6)


def main():
    p = argparse.ArgumentParser(description="Run matches between two scripted bots (batched).")
    p.add_argument("--bot1", choices=BOTS, help="bot type of tank 0")
    p.add_argument("--bot2", choices=BOTS, help="bot type of tank 1")
    p.add_argument("--all", action="store_true", help="every ordered pair of bot types")
    p.add_argument("--num-envs", type=int, default=4096)
    p.add_argument("--ticks", type=int, default=1000)
    p.add_argument("--terminate-time", type=int, default=None, help="default: config_basic.TERMINATE_TIME")
    p.add_argument("--map", default=None, choices=[None, "octagon", "battlefield", "maze"])
    p.add_argument("--seed", type=int, default=0)
    a = p.parse_args()
    if not a.all and not (a.bot1 and a.bot2):
        p.error("--bot1 and --bot2 are required (or --all)")
    pairs = [(b1, b2) for b1 in BOTS for b2 in BOTS] if a.all else [(a.bot1, a.bot2)]
    for b1, b2 in pairs:
        print(_line(run_bot_matches(b1, b2, a.num_envs, a.ticks, "cuda:0", a.seed,
                                    "config" if a.terminate_time is None else a.terminate_time, a.map)), flush=True)


if __name__ == "__main__":
    main()
