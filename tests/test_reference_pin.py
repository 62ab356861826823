"""CPU, build container only: run the UNMODIFIED reference from /root/reference (under the
test-only pygame/gym stubs) side by side with the C oracle.  Skipped where the reference is absent
(the GPU box); tests/test_oracle_golden.py covers the same ground from committed fixtures."""
import numpy as np
import pytest

import parity
import refharness as rh

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.available(), reason="reference sources not present")]


def _live(kind, seed, env_id, ticks, harness=None, **kw):
    rh.load()
    rh.set_config(**{**dict(terminate_time=None, use_octagon=True, buff_on=True, debuff_on=True, elide_aim=True),
                     **(harness or {})})
    kw2 = dict(kw)
    if "game_configs" in kw2:
        kw2["game_configs"] = getattr(rh.load()["mods"]["configs.config_teams"], kw2["game_configs"])
    rec = rh.record_trace(kind, seed, env_id, ticks, **kw2)
    rec["meta"] = dict(kind=kind, seed=seed, env_id=env_id, ticks=ticks, harness=harness or {}, env=kw)
    return rec


@pytest.mark.parametrize("kind,kw,harness", [
    ("agent", {}, {"elide_aim": False}),
    ("bot_agent", {"bot_type": "smart"}, {}),
    ("arena", {"arena_bots": ("dodge", "defensive")}, {"terminate_time": 100}),
    ("team", {"game_configs": "team_vs_bot_configs"}, {}),
])
def test_oracle_tracks_live_reference(kind, kw, harness):
    rec = _live(kind, 1234, 17, 250, harness, **kw)
    env = parity.make_oracle_env(parity.env_spec(rec["meta"]), 1, 1234, 17)
    parity.replay_fixture(env, rec, where="live-" + kind)


def test_aim_raymarch_elision_is_unobservable():
    # F18: Tank._aiming_reward has weight 0; eliding it leaves every output stream identical
    a = _live("team", 5, 1, 150, {"elide_aim": False}, game_configs="team_vs_bot_configs")
    b = _live("team", 5, 1, 150, {"elide_aim": True}, game_configs="team_vs_bot_configs")
    for k in ("obs", "rewards", "done", "tanks", "bullets", "rng_ctr"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("kind,kw", [
    ("agent", {}),
    ("bot_agent", {"bot_type": "aggressive"}),
    ("team", {"game_configs": "team_vs_bot_configs"}),
])
def test_terminal_info_is_the_live_reference_step_info(kind, kw):
    """The info dict the reference's step() returns (gym_env.py:93-98 winner, gym_env_multi.py:218-222 hits / be_hits)
    describes the pre-reset state also at a terminal tick; with the harness's auto-reset the oracle latches it
    (terminal_info) - the behaviour at_get_terminal_info is tested against."""
    from oracle.atrng import synthetic_actions
    rh.load()
    rh.set_config(terminate_time=None, use_octagon=True, buff_on=True, debuff_on=True, elide_aim=True)
    kw2 = dict(kw)
    if "game_configs" in kw2:
        kw2["game_configs"] = getattr(rh.load()["mods"]["configs.config_teams"], kw2["game_configs"])
    seed, env_id, T = 99, 3, 700
    r = rh.RefEnv(kind, seed, env_id, **kw2)
    r.reset()
    meta = dict(kind=kind, seed=seed, env_id=env_id, ticks=T, harness={}, env=kw)
    env = parity.make_oracle_env(parity.env_spec(meta), 1, seed, env_id, auto_reset=True)
    env.reset()
    n_act = r.env.num_agents if kind == "team" else 2
    ended = 0
    for t in range(T):
        acts = synthetic_actions(seed, [env_id], t, n_act)[0]
        obs, rew, done, info, used = r.step(acts)
        ua = np.asarray(used, dtype=np.int32).reshape(-1, 3)[:env.A][None]
        o, rw, d = env.step(ua)
        assert bool(d[0]) == done, t
        assert np.array_equal(np.asarray(rw)[0], rew), t
        ti = env.terminal_info()
        assert int(ti["latched"][0]) == int(done), t
        if kind == "team":
            assert list(ti["hits"][0]) == list(info["hits"].values()), (t, ti["hits"][0], info["hits"])
            assert list(ti["be_hits"][0]) == list(info["be_hits"].values()), t
        else:
            assert int(ti["winner"][0]) == (-1 if info["winner"] is None else info["winner"]), t
        if done:
            ended += 1
            ro, _ = r.reset()
            assert np.array_equal(np.asarray(o)[0], ro), "auto-reset obs differs from the reference's reset() at tick %d" % t
        else:
            assert np.array_equal(np.asarray(o)[0], obs), t
    assert ended >= 2
