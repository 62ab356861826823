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
