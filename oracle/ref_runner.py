"""Times the UNMODIFIED reference env (oracle/_ref/reference, see oracle/make_ref.py) on host cores.

TEST / BENCH INFRASTRUCTURE ONLY - never imported by alphatank_b200/.  Used by `bench.py --impl reference` and by
the `cpu_baseline` leg of the default bench line (BASELINE.md section 3 protocol):

  * the reference's own `env/` + `configs/` packages, imported by path, headless under the repo's pygame / gym
    stand-ins (tests/refstubs; a real pygame is used instead when one is importable - `pygame_kind` says which);
  * `pygame.time.get_ticks` -> virtual clock tick * 1000 // 60; stdout silenced (bfs_path and the observation
    builder print); the render-only GIF decoding of Tank.__init__ / Explosion.__init__ skipped (it would add ~12 ms
    to every reset and has no effect on the simulation - skipping it favours the reference);
  * `_aiming_reward` (the zero-weight ray-march that is 87 % of the reference's time) stays ENABLED, as
    BASELINE.md asks for the headline CPU figure;
  * uniform random agent actions from `random.Random(seed)`, reset on done like the trainers do;
  * (i) one process, one env: `ticks` timed ticks after `warmup` ticks, median of `repeats`;
    (ii) `multiprocessing.Pool(procs)` of independent envs, aggregate env-steps/s.
"""
import contextlib
import importlib
import io
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref", "reference")
STUBS = os.path.join(REPO, "tests", "refstubs")

CONFIGS = {
    1: dict(kind="agent", agents=2, desc="MultiAgentEnv() 1v1, random actions"),
    2: dict(kind="bot_agent", agents=1, desc="MultiAgentEnv(mode='bot_agent', bot_type='smart'), random agent actions"),
    3: dict(kind="team", agents=2, desc="MultiAgentTeamEnv(team_vs_bot_configs): 2 agents vs 2 smart bots"),
    4: dict(kind="arena", agents=2, desc="bot arena defensive vs dodge, TERMINATE_TIME=4096"),
}


def available():
    return os.path.isfile(os.path.join(REF_DIR, "env", "gym_env_multi.py"))


_MODS = {}


def _load():
    if _MODS:
        return _MODS
    assert available(), "oracle/_ref/reference is missing: run `python -m oracle.make_ref` in the build container"
    real_pygame = False
    try:   # a real pygame, if the box has one, is the better reference
        import pygame  # noqa: F401
        real_pygame = not os.path.abspath(getattr(pygame, "__file__", "")).startswith(STUBS)
    except Exception:
        pass
    for p in ((REF_DIR,) if real_pygame else (REF_DIR, STUBS)):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import gym  # noqa: F401
    except Exception:
        if STUBS not in sys.path:
            sys.path.insert(0, STUBS)
    with contextlib.redirect_stdout(io.StringIO()):
        for name in ("pygame", "configs.config_basic", "configs.config_teams", "env.util", "env.sprite", "env.gaming_env",
                     "env.gaming_env_multi", "env.gym_env", "env.gym_env_multi", "env.bots.bot_factory"):
            _MODS[name] = importlib.import_module(name)
    sprite, util = _MODS["env.sprite"], _MODS["env.util"]
    sprite.Tank.load_and_colorize_gif = lambda self, path, color, size: []      # render only

    def _explosion_init(self, x, y, env):
        self.x, self.y, self.env, self.alive = x, y, env, True
        self.frame_index = self.tick = 0
        self.frames, self.total_frames, self.frame_rate, self.played_once = [], 0, 1, False
    util.Explosion.__init__ = _explosion_init
    _MODS["pygame_kind"] = "real" if real_pygame else "stub (tests/refstubs)"
    return _MODS


class _Clock:
    def __init__(self, pg):
        self.pg, self.tick = pg, 0
        self.real = not hasattr(pg.time, "now_ms")
        if self.real:
            pg.time.get_ticks = lambda: self.tick * 1000 // 60

    def advance(self):
        self.tick += 1
        if not self.real:
            self.pg.time.now_ms = self.tick * 1000 // 60


def run_env(config, ticks, warmup, seed=0):
    """One env of BASELINE config `config` for warmup + ticks ticks; returns (seconds of the timed ticks, episodes)."""
    m = _load()
    cfg = CONFIGS[config]
    rng = random.Random(seed)
    random.seed(seed)
    import numpy as np
    np.random.seed(seed)
    clock = _Clock(m["pygame"])
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        if cfg["kind"] == "team":
            env = m["env.gym_env_multi"].MultiAgentTeamEnv(m["configs.config_teams"].team_vs_bot_configs)
            n_act = env.num_agents
        elif cfg["kind"] == "arena":
            for name in ("env.sprite", "env.gym_env", "env.gym_env_multi"):
                m[name].TERMINATE_TIME = 4096
            env = m["env.gym_env"].MultiAgentEnv(mode="bot_vs_bot")
            n_act = 0
        elif cfg["kind"] == "bot_agent":
            env = m["env.gym_env"].MultiAgentEnv(mode="bot_agent", type="train", bot_type="smart", weakness=1.0)
            n_act = 2
        else:
            env = m["env.gym_env"].MultiAgentEnv()
            n_act = 2
        make = m["env.bots.bot_factory"].BotFactory.create_bot
        bots = None
        env.reset()
        if cfg["kind"] == "arena":
            bots = [make(t, env.game_env.tanks[i]) for i, t in enumerate(("defensive", "dodge"))]
        episodes, t0 = 0, 0.0
        for t in range(warmup + ticks):
            if t == warmup:
                t0 = time.perf_counter()
            clock.advance()
            if bots is not None:
                acts = [list(b.get_action()) for b in bots]          # bot_arena.py:51-52
            else:
                acts = [[rng.randrange(3), rng.randrange(3), rng.randrange(2)] for _ in range(n_act)]
            _, _, done, _, _ = env.step(acts)
            if done:
                episodes += 1
                env.reset()
                if bots is not None:
                    bots = [make(tp, env.game_env.tanks[i]) for i, tp in enumerate(("defensive", "dodge"))]
            if sink.tell() > 1 << 20:
                sink.seek(0); sink.truncate()
        dt = time.perf_counter() - t0
    return dt, episodes


def _pool_job(args):
    config, ticks, warmup, seed = args
    dt, _ = run_env(config, ticks, warmup, seed)
    return dt


def time_single(config, ticks=2000, warmup=200, repeats=3):
    """env-steps/s of one env on one core (median of `repeats`)."""
    rates = sorted(ticks / run_env(config, ticks, warmup, seed=r)[0] for r in range(repeats))
    return rates[len(rates) // 2]


def time_pool(config, procs, ticks=600, warmup=50):
    """`procs` independent envs, one per process; aggregate env-steps/s = procs * ticks / wall time of
    the slowest worker's timed region (the workers start together: the pool is created before the clock starts)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        dts = pool.map(_pool_job, [(config, ticks, warmup, 1000 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    return procs * ticks / max(dts), wall


def pygame_kind():
    return _load()["pygame_kind"]


if __name__ == "__main__":
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    r1 = time_single(cfg, ticks=300, warmup=50, repeats=1)
    print("config %d (%s): %.1f env-steps/s on one core [%s]" % (cfg, CONFIGS[cfg]["desc"], r1, pygame_kind()))
    P = os.cpu_count() or 1
    rp, wall = time_pool(cfg, P, ticks=200, warmup=20)
    print("pool of %d: %.1f env-steps/s aggregate (%.1f s)" % (P, rp, wall))
