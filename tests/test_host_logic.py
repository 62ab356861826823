"""CPU: host-side logic of the product - the C ABI library loads and exports every symbol the header declares,
config mirrors equal the reference's, argument validation, and the loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity
import refharness as rh
from alphatank_b200 import _capi, sharding
from alphatank_b200.configs import config_basic, config_teams

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "alphatank_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(at_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    L = _capi.lib()                       # loads without a GPU; no compute call is made here
    for sym in sorted(declared):
        assert hasattr(L, sym), "libalphatank_b200.so does not export %s" % sym
    assert declared == set(_capi.EXPORTS)
    assert b"sm_100a" in L.at_version()


def test_struct_layouts_match_the_header():
    # sizes computed from the header's field lists (LP64)
    assert C.sizeof(_capi.AtConfig) == 8 * 4 + 3 * 8 * 4 + 8
    assert _capi.TANK_DT.itemsize == 3 * 8 + 10 * 4
    assert _capi.BULLET_DT.itemsize == 5 * 8 + 4 * 4
    assert _capi.BOT_DT.itemsize == 8 * 4 + 16 + 8 + 16 + 16 + 8 + 16
    from oracle import oracle as O
    assert _capi.ENV_STATE_DT.itemsize == O.ENV_DT.itemsize     # oracle and product exchange the same snapshot record


def test_policy_struct_layout_and_argument_checks():
    """at_policy (include/alphatank_b200.h) as the ctypes mirror lays it out, and the argument checks of the policy entry
    points, which come before any CUDA call (so they can be exercised without a GPU)."""
    P = _capi.AtPolicy
    assert C.sizeof(P) == 16 * 8                       # 15 pointer / int64 fields + uint32 stream_id, padded (LP64)
    hdr = open(os.path.join(REPO, "include", "alphatank_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr[hdr.index("typedef struct at_policy {"):hdr.index("} at_policy;")], flags=re.S)
    names = re.findall(r"[\*\s](\w+)\s*[;,]", body)
    assert names == [n for n, _ in P._fields_], names
    L = _capi.lib()
    for fn in (L.at_policy_forward, L.at_policy_forward16):
        assert fn(0, None, 1, 128, 528, 0, 0, None, 0, None) == _capi.AT_ERR_ARG            # no policies
        pol = (P * 1)()
        assert fn(0, pol, 3, 128, 528, 0, 0, None, 0, None) == _capi.AT_ERR_ARG             # more than two
        assert fn(0, pol, 1, 0, 528, 0, 0, None, 0, None) == _capi.AT_ERR_ARG               # no rows
        assert fn(0, pol, 1, 128, 530, 0, 0, None, 0, None) == _capi.AT_ERR_ARG             # row length not a multiple of 4 / 8
        assert fn(0, pol, 1, 128, 528, 1 << 20, 0, None, 0, None) == _capi.AT_ERR_ARG       # chunk past the last column
    assert L.at_policy_forward16(0, (P * 1)(), 1, 128, 388, 0, 0, None, 0, None) == _capi.AT_ERR_ARG   # 388 % 8 != 0: fp32 rows only
    assert b"at_policy_forward" in L.at_last_error()
    assert L.at_obs_static_range(None, None, None) == _capi.AT_ERR_ARG
    assert L.at_set_rows_mode(None, 1) == _capi.AT_ERR_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv
    with pytest.raises(RuntimeError):
        MultiAgentEnv(num_envs=2)
    with pytest.raises(RuntimeError):
        MultiAgentTeamEnv(config_teams.team_vs_bot_configs, num_envs=2)
    h = C.c_void_p()
    cfg = _capi.make_config("agent", [0, 1], ["agent", "agent"], [None, None])
    rc = _capi.lib().at_create(C.byref(cfg), 4, 0, 0, 0, C.byref(h))
    assert rc == _capi.AT_ERR_NOGPU and b"no CPU fallback" in _capi.lib().at_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "alphatank_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                if "cpu_emul" in src:      # the shared headers may only MENTION the test-only emulator in comments
                    assert f in ("at_sim.h", "at_host.h", "at_rows.h", "at_warp.h"), f
                    for line in src.splitlines():
                        assert "cpu_emul" not in line or line.lstrip().startswith("//"), (f, line)


def test_argument_validation():
    with pytest.raises(ValueError):
        _capi.make_config("team", [0], ["agent"], [None])
    with pytest.raises(ValueError):
        _capi.make_config("bot_agent", [0, 1], ["bot", "agent"], ["nonsense", None])
    with pytest.raises(ValueError):
        _capi.make_config("agent", [0, 1], ["agent", "agent"], [None, None], map="desert")
    with pytest.raises(KeyError):
        _capi.make_config("human_play", [0, 1], ["agent", "agent"], [None, None])


@pytest.mark.skipif(not rh.available(), reason="reference sources not present")
def test_config_mirrors_equal_the_reference():
    m = rh.load()["mods"]
    rb, rt = m["configs.config_basic"], m["configs.config_teams"]
    for name in dir(rb):
        if name.isupper() and not name.startswith("K_"):
            assert getattr(config_basic, name) == getattr(rb, name), name
    for name in ("team_configs", "pair_configs", "team_vs_bot_configs", "team_vs_bot_hard_configs",
                 "team_vs_bot_harder_configs", "team_vs_def_bot_configs", "bot_team_configs", "mixed_team_configs"):
        ours, ref = getattr(config_teams, name), getattr(rt, name)
        assert list(ours) == list(ref), name
        for k in ref:
            for f in ("team", "color", "mode", "bot_type"):
                assert ours[k].get(f) == ref[k].get(f), (name, k, f)
    assert config_teams.inference_agent_configs == rt.inference_agent_configs
    # the layouts the tests use are the same dictionaries
    for name, layout in parity.TEAM_DICTS.items():
        ref = getattr(rt, name)
        assert [(v["team"][-1], v["mode"], v.get("bot_type")) for v in ref.values()] == layout, name


def test_shard_ranges_tile_the_id_space():
    for total, w in ((1048576, 8), (65536, 1), (10, 3), (7, 8)):
        spans = [sharding.shard_range(total, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
