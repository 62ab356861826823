"""CPU: the C oracle replays every golden trace recorded from the unmodified reference.
Bit-exact on events, integer state, f64 positions/rewards, f32 obs, and the RNG draw count."""
import numpy as np
import pytest

import parity
from oracle import oracle as O


@pytest.mark.parametrize("name", parity.fixture_names())
def test_oracle_replays_reference_trace(name):
    rec = parity.load_fixture(name)
    meta = rec["meta"]
    env = parity.make_oracle_env(parity.env_spec(meta), 1, meta["seed"], meta["env_id"])
    assert env.D * env.R == rec["obs"].shape[1]
    ticks = parity.replay_fixture(env, rec, where=name)
    assert ticks == meta["ticks"]


def test_obs_dims_match_shipped_checkpoint_widths():
    # first-layer widths of the reference's demo checkpoints
    def d(mode, layout, map_):
        spec = dict(mode=mode, teams=[0 if t == "A" else 1 for t, _, _ in layout], ctrl=[c for _, c, _ in layout],
                    bots=[b for _, _, b in layout], map=map_, terminate_time=None, buff_on=True, debuff_on=True,
                    weakness=1.0)
        return parity.make_oracle_env(spec, 1, 0).D
    two = [("A", "agent", None), ("B", "agent", None)]
    assert d("agent", two, "octagon") == 388
    assert d("agent", two, "battlefield") == 472
    T = parity.TEAM_DICTS
    assert d("team", T["team_vs_def_bot_configs"], "octagon") == 462
    assert d("team", T["team_vs_bot_configs"], "octagon") == 528
    assert d("team", T["team_vs_bot_configs"], "battlefield") == 612
    assert d("team", T["team_vs_bot_hard_configs"], "octagon") == 594
    assert d("team", T["team_vs_bot_harder_configs"], "battlefield") == 744


def test_auto_reset_equals_manual_reset():
    rec = parity.load_fixture("team_2v2_vs_bot")
    meta = rec["meta"]
    spec = parity.env_spec(meta)
    a = parity.make_oracle_env(spec, 1, meta["seed"], meta["env_id"], auto_reset=True)
    a.reset()
    ri = 1
    for t in range(600):
        obs, rew, done = a.step(rec["actions"][t][:a.A][None])
        assert done[0] == rec["done"][t]
        assert np.array_equal(rew[0], rec["rewards"][t])       # terminal rewards survive the in-step reset
        if done[0]:
            assert np.array_equal(obs[0], rec["reset_obs"][ri])  # obs is the post-reset observation
            ri += 1
        else:
            assert np.array_equal(obs[0], rec["obs"][t])
    assert ri > 1


def test_batch_equals_singles_and_threads():
    # env i of a batch == a single env created with env_id_base = i (shard invariance), any thread count
    spec = parity.env_spec(parity.load_fixture("team_2v2_vs_bot")["meta"])
    from oracle.atrng import synthetic_actions
    N, T = 16, 120
    batch = parity.make_oracle_env(spec, N, 99, 0, auto_reset=True)
    mt = parity.make_oracle_env(spec, N, 99, 0, auto_reset=True)
    singles = [parity.make_oracle_env(spec, 1, 99, i, auto_reset=True) for i in range(N)]
    batch.reset(); mt.reset()
    for s in singles:
        s.reset()
    for t in range(T):
        acts = synthetic_actions(99, np.arange(N), t, batch.A)
        ob, rb, db = batch.step(acts)
        om, rm, dm = mt.step(acts, threads=3)
        assert np.array_equal(ob, om) and np.array_equal(rb, rm) and np.array_equal(db, dm)
        for i, s in enumerate(singles):
            o1, r1, d1 = s.step(acts[i:i + 1])
            assert np.array_equal(o1[0], ob[i]) and np.array_equal(r1[0], rb[i]) and d1[0] == db[i]
