"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference) under tests/refharness.py.  Run in the build container:

    python tests/golden/gen_golden.py            # all fixtures
    python tests/golden/gen_golden.py team_2v2   # one fixture

Each fixture is one reference env (one ATRNG seed / env id) stepped with the
synthetic action stream (oracle/atrng.py, stream 1+a) and the trainers'
reset-on-done policy.  Arrays (T = ticks, n = tanks, R*D = flat obs):

    obs[T,R*D] f32        what env.step() returned            (post-step, pre-reset)
    rewards[T,R] f32      cumulative rewards as returned (F1)
    done[T] u8
    tanks[T,n,12] f64     x y angle speed alive hittingWall reward num_hit num_be_hit in_buff in_debuff max_bullets
    nbul[T] i32, bullets[T,6n,9] f64  owner x y dx dy distance bounces pending cleared (env.bullets order)
    actions[T,A,3] i32    actions fed (bots' own actions in arena mode)
    rng_ctr[T] i64        ATRNG draws consumed so far
    reset_obs[E,R*D], reset_tick[E], reset_tanks[E,n,12], reset_zones[E,8], maze[11,11]
    meta                  json: kind, seed, env_id, ticks and the config kwargs

The fixtures are the reference's outputs; tests/test_oracle_golden.py replays
them through the C oracle and (on a GPU) tests/test_gpu_parity.py through the
CUDA path.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402

# name -> (kind, seed, env_id, ticks, harness config, RefEnv kwargs)
FIXTURES = {
    # BASELINE config 1: 1v1, random actions, octagon.  aim ray-march NOT elided (fully faithful run)
    "agent_1v1_faithful": ("agent", 11, 0, 500, dict(elide_aim=False), {}),
    "agent_1v1": ("agent", 12, 5, 1600, dict(), {}),
    "agent_1v1_terminate128": ("agent", 13, 2, 900, dict(terminate_time=128), {}),
    "agent_1v1_battlefield": ("agent", 14, 1, 1200, dict(use_octagon=False), {}),
    "agent_1v1_nobuff": ("agent", 15, 4, 600, dict(buff_on=False), {}),
    # BASELINE config 2: ppo-vs-bot workload
    "bot_agent_smart": ("bot_agent", 21, 7, 1500, dict(), dict(bot_type="smart")),
    "bot_agent_smart_battlefield": ("bot_agent", 22, 3, 1000, dict(use_octagon=False), dict(bot_type="smart")),
    "bot_agent_aggressive_weak": ("bot_agent", 23, 9, 1000, dict(), dict(bot_type="aggressive", weakness=0.7)),
    "bot_agent_defensive": ("bot_agent", 24, 1, 800, dict(), dict(bot_type="defensive")),
    "bot_agent_dodge": ("bot_agent", 25, 2, 800, dict(), dict(bot_type="dodge")),
    "bot_agent_random": ("bot_agent", 26, 6, 800, dict(), dict(bot_type="random", weakness=0.5)),
    "bot_agent_stationary": ("bot_agent", 27, 8, 400, dict(), dict(bot_type="stationary")),
    # BASELINE config 4: arena stress (immortal tanks, F6 victory-bonus-on-every-hit quirk)
    "arena_def_dodge_t256": ("arena", 31, 0, 1200, dict(terminate_time=256), dict(arena_bots=("defensive", "dodge"))),
    "arena_dodge_aggr": ("arena", 32, 4, 1000, dict(), dict(arena_bots=("dodge", "aggressive"))),
    "arena_smart_smart_battlefield": ("arena", 33, 5, 800, dict(use_octagon=False),
                                      dict(arena_bots=("smart", "smart"))),
    # BASELINE configs 3/5: team mode
    "team_2v2_vs_bot": ("team", 41, 10, 1500, dict(), dict(game_configs="team_vs_bot_configs")),
    "team_2v2_vs_bot_b": ("team", 42, 65535, 1000, dict(), dict(game_configs="team_vs_bot_configs")),
    "team_2v2_agents": ("team", 43, 1, 500, dict(), dict(game_configs="team_configs")),
    "team_pair": ("team", 44, 2, 600, dict(), dict(game_configs="pair_configs")),
    "team_vs_bot_hard": ("team", 45, 3, 800, dict(), dict(game_configs="team_vs_bot_hard_configs")),
    "team_vs_bot_harder_battlefield": ("team", 46, 4, 600, dict(use_octagon=False),
                                       dict(game_configs="team_vs_bot_harder_configs")),
    "team_vs_def_bot": ("team", 47, 5, 600, dict(), dict(game_configs="team_vs_def_bot_configs")),
    "team_three_bot_teams": ("team", 48, 6, 800, dict(), dict(game_configs="bot_team_configs")),
    "team_2v2_vs_bot_t64": ("team", 49, 7, 500, dict(terminate_time=64), dict(game_configs="team_vs_bot_configs")),
}


def make(name):
    kind, seed, env_id, ticks, hcfg, kw = FIXTURES[name]
    rh.load()
    rh.set_config(**{**dict(terminate_time=None, use_octagon=True, buff_on=True, debuff_on=True, elide_aim=True),
                     **hcfg})
    kw2 = dict(kw)
    if "game_configs" in kw2:
        kw2["game_configs"] = getattr(rh.load()["mods"]["configs.config_teams"], kw2["game_configs"])
    t0 = time.time()
    rec = rh.record_trace(kind, seed, env_id, ticks, **kw2)
    meta = dict(kind=kind, seed=seed, env_id=env_id, ticks=ticks, harness=hcfg, env=kw)
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-36s %5d ticks %3d episodes maxbul %2d  %.0fs  %.0f KB" % (
        name, ticks, int(rec["done"].sum()), int(rec["nbul"].max()), time.time() - t0, os.path.getsize(path) / 1024))


def make_mazes(n=64):
    """env/maze.py:7-33 generate_maze(11, 11) with every draw served by ATRNG (seed 5, env id = index)."""
    from oracle.atrng import Stream
    st = rh.load()
    gm = st["mods"]["env.maze"].generate_maze
    mazes, ctrs = [], []
    for i in range(n):
        rh.RANDOM.stream = Stream(5, i)
        mazes.append(np.asarray(gm(11, 11), dtype=np.uint8))
        ctrs.append(rh.RANDOM.stream.ctr)
    np.savez_compressed(os.path.join(HERE, "mazes.npz"), mazes=np.asarray(mazes), ctrs=np.asarray(ctrs), seed=5)
    print("mazes: %d, open cells %s" % (n, sorted(set(int((m == 0).sum()) for m in mazes))))


def make_bfs(n_mazes=24):
    """env/bfs.py:6-56 bfs_path on octagon, battlefield and generated mazes: every free (start, goal) pair's
    len(path) and path[1] (the only parts of the path the env consumes)."""
    import contextlib
    import io
    from oracle.atrng import Stream
    st = rh.load()
    gm = st["mods"]["env.maze"].generate_maze
    bfs = st["mods"]["env.bfs"].bfs_path
    grids = []
    oct_ = np.ones((11, 11), np.uint8); oct_[1:-1, 1:-1] = 0
    grids.append(oct_)
    rh.set_config(use_octagon=False)
    g = st["mods"]["env.gaming_env"].GamingENV.__new__(st["mods"]["env.gaming_env"].GamingENV)
    g.GRID_SIZE = 70.0
    g.constructWall()
    grids.append(np.asarray(g.maze, dtype=np.uint8))
    rh.set_config()
    for i in range(n_mazes):
        rh.RANDOM.stream = Stream(6, i)
        grids.append(np.asarray(gm(11, 11), dtype=np.uint8))
    recs = []
    with contextlib.redirect_stdout(io.StringIO()):
        for gi, grid in enumerate(grids):
            gl = grid.tolist()
            free = [(r, c) for r in range(11) for c in range(11) if grid[r, c] == 0]
            for s in free:
                for t in free:
                    p = bfs(gl, s, t)
                    nxt = p[1] if p is not None and len(p) > 1 else (-1, -1)
                    recs.append((gi, s[0], s[1], t[0], t[1], 0 if p is None else len(p), nxt[0], nxt[1]))
    np.savez_compressed(os.path.join(HERE, "bfs.npz"), grids=np.asarray(grids), cases=np.asarray(recs, np.int16))
    print("bfs: %d grids, %d (start, goal) cases" % (len(grids), len(recs)))


if __name__ == "__main__":
    names = sys.argv[1:] or (list(FIXTURES) + ["mazes", "bfs"])
    for nm in names:
        if nm == "mazes":
            make_mazes()
        elif nm == "bfs":
            make_bfs()
        else:
            make(nm)
