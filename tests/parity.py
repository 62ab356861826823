"""Shared replay/compare helpers for the parity tests (test infrastructure).

``replay_fixture`` feeds a golden trace (recorded from the unmodified reference,
tests/golden/gen_golden.py) through any env object exposing the small protocol
below and compares every tick:

    env.reset() -> obs[N, R*D]             env.step(actions) -> (obs, rewards[N,R], done[N])
    env.get_state(i) -> ENV_DT record      env.A, env.n_tanks

Both oracle.oracle.OracleEnv (CPU) and tests/gpu_env.GpuEnv (CUDA C-ABI) implement it.
Bar: bit-exact for events / integer state / positions (f64) / obs (f32);
rewards exact unless ``reward_rtol`` is given.
"""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TANK_COLS = (("x", 0), ("y", 1), ("angle", 2), ("speed", 3), ("alive", 4), ("hitting_wall", 5), ("num_hit", 7),
             ("num_be_hit", 8), ("in_buff", 9), ("in_debuff", 10), ("max_bullets", 11))
BULLET_COLS = (("owner", 0), ("x", 1), ("y", 2), ("dx", 3), ("dy", 4), ("distance", 5), ("bounces", 6),
               ("pending", 7), ("cleared", 8))


def fixture_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f not in ("mazes.npz", "bfs.npz", "learner.npz"))


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    rec["meta"] = json.loads(bytes(rec["meta"]).decode())
    return rec


TEAM_DICTS = {
    # team/ctrl/bot layout of configs/config_teams.py (names only; no pygame needed)
    "team_configs": [("A", "agent", None)] * 2 + [("B", "agent", None)] * 2,
    "pair_configs": [("A", "agent", None), ("B", "agent", None)],
    "team_vs_bot_configs": [("A", "agent", None)] * 2 + [("B", "bot", "smart")] * 2,
    "team_vs_bot_hard_configs": [("A", "agent", None)] * 2 + [("B", "bot", "aggressive"), ("B", "bot", "smart"),
                                                              ("B", "bot", "defensive")],
    "team_vs_bot_harder_configs": [("A", "agent", None)] * 2 + [("B", "bot", "aggressive"), ("B", "bot", "smart"),
                                                                ("B", "bot", "defensive"), ("B", "bot", "smart")],
    "team_vs_def_bot_configs": [("A", "agent", None)] * 2 + [("B", "bot", "defensive")],
    "bot_team_configs": [("A", "bot", "smart")] * 2 + [("B", "bot", "aggressive")] * 2 + [("C", "bot", "defensive")] * 2,
}


def env_spec(meta):
    """Fixture meta -> generic spec dict understood by make_oracle_env / gpu_env.make_gpu_env."""
    h, e = meta["harness"], meta["env"]
    spec = dict(map="octagon" if h.get("use_octagon", True) else "battlefield",
                terminate_time=h.get("terminate_time"), buff_on=h.get("buff_on", True),
                debuff_on=h.get("debuff_on", True), weakness=e.get("weakness", 1.0))
    kind = meta["kind"]
    if kind == "agent":
        spec.update(mode="agent", teams=[0, 1], ctrl=["agent", "agent"], bots=[None, None])
    elif kind == "bot_agent":
        spec.update(mode="bot_agent", teams=[0, 1], ctrl=["bot", "agent"], bots=[e["bot_type"], None])
    elif kind == "arena":
        spec.update(mode="arena", teams=[0, 1], ctrl=["bot", "bot"], bots=list(e["arena_bots"]))
    else:
        layout = TEAM_DICTS[e["game_configs"]]
        ids = {}
        spec.update(mode="team", teams=[ids.setdefault(t, len(ids)) for t, _, _ in layout],
                    ctrl=[c for _, c, _ in layout], bots=[b for _, _, b in layout])
    return spec


def make_oracle_env(spec, n_envs, seed, env_id_base=0, auto_reset=False):
    from oracle import oracle as O
    cfg = O.make_config(spec["mode"], spec["teams"], spec["ctrl"], spec["bots"], map=spec["map"],
                        terminate_time=spec["terminate_time"], buff_on=spec["buff_on"], debuff_on=spec["debuff_on"],
                        auto_reset=auto_reset, weakness=spec["weakness"])
    return O.OracleEnv(cfg, n_envs, seed, env_id_base)


class Mismatch(AssertionError):
    pass


def _chk(cond, msg):
    if not cond:
        raise Mismatch(msg)


def compare_state_to_trace(st, rec, t, n, where, reward_rtol=None):
    tk = rec["tanks"][t]
    for name, col in TANK_COLS:
        got = st["tanks"][name][:n].astype(np.float64)
        _chk(np.array_equal(got, tk[:, col]), "%s tick %d tank.%s got %s want %s" % (where, t, name, got, tk[:, col]))
    nb = int(st["n_bullets"])
    _chk(nb == rec["nbul"][t], "%s tick %d n_bullets got %d want %d" % (where, t, nb, rec["nbul"][t]))
    for name, col in BULLET_COLS:
        got = st["bullets"][name][:nb].astype(np.float64)
        want = rec["bullets"][t][:nb, col]
        _chk(np.array_equal(got, want), "%s tick %d bullet.%s got %s want %s" % (where, t, name, got, want))
    _chk(int(st["rng_ctr"]) == rec["rng_ctr"][t], "%s tick %d rng_ctr got %d want %d" % (
        where, t, st["rng_ctr"], rec["rng_ctr"][t]))
    got = st["tanks"]["reward"][:n]
    if reward_rtol is None:
        _chk(np.array_equal(got, tk[:, 6]), "%s tick %d tank.reward(f64) got %r want %r" % (where, t, got, tk[:, 6]))
    else:
        _chk(np.allclose(got, tk[:, 6], rtol=reward_rtol, atol=reward_rtol),
             "%s tick %d tank.reward(f64) got %r want %r" % (where, t, got, tk[:, 6]))


def replay_fixture(env, rec, where="", max_ticks=None, reward_rtol=None, check_state_every=1):
    """env must have been created with the fixture's (seed, env_id) and n_envs == 1."""
    n = rec["tanks"].shape[1]
    T = rec["obs"].shape[0] if max_ticks is None else min(max_ticks, rec["obs"].shape[0])
    obs0 = np.asarray(env.reset())[0]
    _chk(np.array_equal(obs0, rec["reset_obs"][0]), where + " initial reset obs differs")
    ri = 1
    for t in range(T):
        acts = rec["actions"][t][:env.A][None] if env.A else None
        obs, rew, done = env.step(acts)
        obs, rew, done = np.asarray(obs), np.asarray(rew), np.asarray(done)
        _chk(int(done[0]) == int(rec["done"][t]), "%s tick %d done got %d want %d" % (where, t, done[0], rec["done"][t]))
        if t % check_state_every == 0 or done[0]:
            compare_state_to_trace(env.get_state(0), rec, t, n, where, reward_rtol)
        if reward_rtol is None:
            _chk(np.array_equal(rew[0], rec["rewards"][t]), "%s tick %d rewards got %r want %r" % (
                where, t, rew[0], rec["rewards"][t]))
        else:
            _chk(np.allclose(rew[0], rec["rewards"][t], rtol=reward_rtol, atol=reward_rtol),
                 "%s tick %d rewards got %r want %r" % (where, t, rew[0], rec["rewards"][t]))
        if not np.array_equal(obs[0], rec["obs"][t]):
            d = np.nonzero(obs[0] != rec["obs"][t])[0]
            raise Mismatch("%s tick %d obs differs at cols %s got %s want %s" % (
                where, t, d[:12], obs[0][d[:6]], rec["obs"][t][d[:6]]))
        if done[0]:
            ob = np.asarray(env.reset())[0]
            _chk(np.array_equal(ob, rec["reset_obs"][ri]), "%s tick %d post-done reset obs differs" % (where, t))
            ri += 1
    return T
