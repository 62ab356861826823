"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes -> libalphatank_b200.so), against
 (a) the golden traces recorded from the unmodified reference,
 (b) the CPU oracle on seeded batches, and
 (c) size-independent properties at BASELINE.json's full batch sizes.
Bar: bit-exact events / integer state / f64 positions / f32 obs; rewards bit-exact as well (the kernels
restate the reference's fp64 operation order, including numpy's FMA-contracted 2-vector dots)."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle.atrng import synthetic_actions
from test_kernel_logic_cpu import SWEEP

pytestmark = pytest.mark.gpu


def _gpu(*a, **k):
    from gpu_env import GpuEnv
    return GpuEnv(*a, **k)


def test_native_library_is_the_cuda_one():
    from alphatank_b200 import _capi
    L = _capi.lib()
    assert b"sm_100a" in L.at_version()
    assert torch.cuda.get_device_capability(0)[0] >= 10


@pytest.mark.parametrize("name", parity.fixture_names())
def test_cuda_replays_reference_trace(name):
    rec = parity.load_fixture(name)
    meta = rec["meta"]
    env = _gpu(parity.env_spec(meta), 1, meta["seed"], meta["env_id"])
    assert env.D * env.R == rec["obs"].shape[1]
    parity.replay_fixture(env, rec, where=name, check_state_every=2)


@pytest.mark.parametrize("name", sorted(SWEEP))
def test_cuda_matches_oracle_on_batches(name):
    spec = SWEEP[name]
    N, ticks, seed, base = 1000, 220, 77, 5000          # 1000: last block partially filled
    a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
    b = _gpu(spec, N, seed, base, auto_reset=True)
    assert (a.D, a.R, a.A) == (b.D, b.R, b.A)
    assert np.array_equal(a.reset(), b.reset())
    ids = np.arange(base, base + N)
    thr = os.cpu_count() or 1
    dones = 0
    for t in range(ticks):
        acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
        oa, ra, da = a.step(acts, threads=thr)
        ob, rb, db = b.step(acts)
        assert np.array_equal(da, db), "%s tick %d done differs for envs %s" % (name, t, np.nonzero(da != db)[0][:8])
        assert np.array_equal(ra, rb), "%s tick %d rewards differ for envs %s" % (
            name, t, np.unique(np.nonzero(ra != rb)[0])[:8])
        if not np.array_equal(oa, ob):
            raise AssertionError("%s tick %d obs differ at (env, col) %s" % (name, t, np.argwhere(oa != ob)[:8].tolist()))
        if da.any() or t == 0:   # info of the step as the reference's callers see it: the finished episode's, latched before the auto-reset
            ta, tb = a.terminal_info(), b.terminal_info()
            for f in ta:
                assert np.array_equal(ta[f], tb[f]), "%s tick %d terminal info %s" % (name, t, f)
            assert np.array_equal(tb["latched"], db)
        dones += int(da.sum())
    sa, sb = a.get_state(), b.get_state()
    n = a.n_tanks
    for f in ("n_bullets", "tick", "training_step", "episode_steps", "rng_ctr", "zones", "maze"):
        assert np.array_equal(sa[f], sb[f]), f
    for f in ("x", "y", "reward", "angle", "speed", "alive", "hitting_wall", "in_buff", "in_debuff", "max_bullets",
              "last_shot_ms", "num_hit", "num_be_hit"):
        assert np.array_equal(sa["tanks"][f][:, :n], sb["tanks"][f][:, :n]), f
    ia, ib = a.info(), b.info()
    for f in ia:
        assert np.array_equal(ia[f], ib[f]), f
    assert dones > 0 or spec["mode"] == "agent"


def test_device_action_stream_is_atrng():
    b = _gpu(SWEEP["team_vs_bot"], 300, 9, 123)
    for tick in (0, 1, 77, 2**20):
        got = b.core.synthetic_actions(tick).cpu().numpy()
        assert np.array_equal(got, synthetic_actions(9, np.arange(123, 423), tick, b.A))


def test_bfs_kernel_matches_reference_pairs():
    from alphatank_b200 import _capi
    import ctypes as C
    z = np.load(os.path.join(parity.GOLDEN, "bfs.npz"))
    grids = torch.as_tensor(z["grids"].reshape(len(z["grids"]), -1).astype(np.uint8)).cuda()
    cases = z["cases"].astype(np.int32)
    q = torch.as_tensor(np.ascontiguousarray(cases[:, :5])).cuda()
    out = torch.zeros((len(cases), 3), dtype=torch.int32, device="cuda")
    _capi.check(_capi.lib().at_bfs_batch(0, C.c_void_p(grids.data_ptr()), C.c_void_p(q.data_ptr()), len(cases),
                                         C.c_void_p(out.data_ptr()), None), "at_bfs_batch")
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.array_equal(o[:, 0], cases[:, 5])                       # len(path): BFS distances bit-exact
    moving = cases[:, 5] > 1
    assert np.array_equal(o[moving, 1:], cases[moving, 6:8])          # path[1], incl. the FIFO tie-break


def test_maze_kernel_matches_reference():
    from alphatank_b200 import _capi
    import ctypes as C
    z = np.load(os.path.join(parity.GOLDEN, "mazes.npz"))
    n = len(z["mazes"])
    m = torch.zeros((n, 121), dtype=torch.uint8, device="cuda")
    c = torch.zeros((n,), dtype=torch.int32, device="cuda")
    _capi.check(_capi.lib().at_generate_mazes(0, int(z["seed"]), 0, n, C.c_void_p(m.data_ptr()),
                                              C.c_void_p(c.data_ptr()), None), "at_generate_mazes")
    torch.cuda.synchronize()
    assert np.array_equal(m.cpu().numpy().reshape(n, 11, 11), z["mazes"])
    assert np.array_equal(c.cpu().numpy(), z["ctrs"])


def test_lookup_tables_are_libm():
    import math
    b = _gpu(SWEEP["agent_t96"], 1, 0)
    trig, corn = b.core.luts()
    for a in range(361):
        assert trig[a, 0] == math.cos(math.radians(a)) and trig[a, 1] == math.sin(math.radians(a))
    from oracle import oracle as O
    out = np.zeros(8)
    for a in range(361):
        O.lib().ato_corners(0.0, 0.0, a, out.ctypes.data)
        assert np.array_equal(corn[a], out[:4])


# ----------------------------------------------------------------- BASELINE.json sizes: properties
def _full(name, N, seed=3, base=0, **kw):
    return _gpu(SWEEP[name], N, seed, base, auto_reset=True, **kw)


def test_full_size_2v2_shard_invariance_determinism_and_oracle_slices():
    """config 3 (65536 envs, 2 agents vs 2 smart bots): (i) two identical runs are bitwise equal, (ii) env i of the
    big batch equals the same env id stepped in a small batch (how the 8-GPU sharding cuts the id space),
    (iii) the first / last envs equal the CPU oracle, (iv) invariants on the whole batch."""
    N, ticks, seed = 65536, 160, 3
    big = _full("team_vs_bot", N, seed)
    big2 = _full("team_vs_bot", N, seed)
    lo = _full("team_vs_bot", 96, seed, 0)
    hi = _full("team_vs_bot", 96, seed, N - 96)
    ora = parity.make_oracle_env(SWEEP["team_vs_bot"], 96, seed, N - 96, auto_reset=True)
    for e in (big, big2, lo, hi, ora):
        e.reset()
    dones = np.zeros(N, np.int64)
    acts = torch.zeros((N, big.A, 3), dtype=torch.int32, device="cuda")
    for t in range(ticks):
        big.core.synthetic_actions(t, acts)
        o1, r1, d1 = big.core.step(acts)
        o1, r1, d1 = o1.clone(), r1.clone(), d1.clone()
        o2, r2, d2 = big2.core.step(acts)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), "run-to-run determinism, tick %d" % t
        a = acts.cpu().numpy()
        ol, rl, dl = lo.step(a[:96])
        oh, rh, dh = hi.step(a[-96:])
        oo, ro, do = ora.step(a[-96:], threads=4)
        o1n, r1n, d1n = o1.cpu().numpy(), r1.cpu().numpy(), d1.cpu().numpy()
        assert np.array_equal(o1n[:96], ol) and np.array_equal(r1n[:96], rl) and np.array_equal(d1n[:96], dl)
        assert np.array_equal(o1n[-96:], oh) and np.array_equal(r1n[-96:], rh) and np.array_equal(d1n[-96:], dh)
        assert np.array_equal(oh, oo) and np.array_equal(rh, ro) and np.array_equal(dh, do), "oracle slice, tick %d" % t
        assert np.isfinite(o1n).all() and o1n.min() >= -771.0 and o1n.max() <= 900.0   # dist <= 630*sqrt(2)
        dones += d1n
    st = big.get_state()
    assert (st["n_bullets"] <= 24).all() and (st["n_bullets"] >= 0).all()
    assert (st["tick"] == ticks).all()
    tk = st["tanks"]
    assert ((tk["x"][:, :4] > 70) & (tk["x"][:, :4] < 700) & (tk["y"][:, :4] > 70) & (tk["y"][:, :4] < 700)).all()
    assert ((tk["angle"][:, :4] >= 0) & (tk["angle"][:, :4] <= 360)).all()
    assert dones.sum() > N // 4                                      # episodes do end (about 190 ticks each)
    assert (tk["num_hit"][:, :4].sum(1) == tk["num_be_hit"][:, :4].sum(1)).all()   # every hit has a victim


def test_full_size_arena_stress_bullet_saturation():
    """config 4 (262144 envs, defensive vs dodge, TERMINATE_TIME=4096 -> immortal tanks, bullets saturate)."""
    N = 262144
    spec = dict(SWEEP["arena_def_dodge_t200"], terminate_time=4096)
    env = _gpu(spec, N, 11, 0, auto_reset=True)
    env.core.reset()
    for t in range(120):
        env.core.step(None)
    st = env.get_state(0, 4096) if False else env.core.get_state(0, 4096)
    assert st["n_bullets"].max() == 12 or st["n_bullets"].max() >= 10
    assert (st["tanks"]["alive"][:, :2] == 1).all()                  # nobody dies with TERMINATE_TIME set
    d = env.core.done.cpu().numpy()
    assert d.sum() == 0                                              # 120 < 4096
    ora = parity.make_oracle_env(spec, 64, 11, 0, auto_reset=True)
    ora.reset()
    for t in range(120):
        ora.step(None)
    so = ora.get_state()
    assert np.array_equal(so["tanks"]["reward"][:, :2], st["tanks"]["reward"][:64, :2])
    assert np.array_equal(so["n_bullets"], st["n_bullets"][:64])


@pytest.mark.parametrize("dtype, pinned", [(torch.int32, True), (torch.uint8, True), (torch.uint8, False)])
def test_step_host_equals_step_and_non_default_stream(dtype, pinned):
    """at_step_host (int32 host actions) and at_step_host_u8 (one byte per action, widened on the device; pinned = the
    DMA + mapped-results path, pageable = the copy path) against at_step, on a non-default stream."""
    spec = SWEEP["team_vs_bot"]
    a = _gpu(spec, 512, 5, 0, auto_reset=True)
    b = _gpu(spec, 512, 5, 0, auto_reset=True)
    a.reset(); b.reset()
    s = torch.cuda.Stream()
    h = torch.zeros((512, a.A, 3), dtype=dtype)
    if pinned:
        h = h.pin_memory()
    for t in range(60):
        acts = synthetic_actions(5, np.arange(512), t, a.A)
        h.copy_(torch.as_tensor(acts))
        with torch.cuda.stream(s):
            ob, rb, db = b.core.step_host(h)
        oa, ra, da = a.step(acts)
        assert np.array_equal(oa, ob.cpu().numpy()) and np.array_equal(ra, rb.numpy()) and np.array_equal(da, db.numpy())


def test_masked_reset_only_touches_masked_envs():
    spec = SWEEP["agent_t96"]
    env = _gpu(spec, 200, 8, 0)
    env.reset()
    for t in range(10):
        env.step(synthetic_actions(8, np.arange(200), t, env.A))
    before = env.get_state()
    obs_before = env.core.obs.cpu().numpy().copy()
    mask = np.zeros(200, np.uint8)
    mask[[3, 64, 199]] = 1
    obs = env.reset(mask)
    after = env.get_state()
    keep = mask == 0
    assert np.array_equal(before[keep], after[keep])
    assert np.array_equal(obs[keep], obs_before[keep])
    assert (after["tanks"]["reward"][mask == 1][:, :2] == 0).all() and (after["n_bullets"][mask == 1] == 0).all()
    assert (after["rng_ctr"][mask == 1] > before["rng_ctr"][mask == 1]).all()


def test_reference_facing_classes_and_errors():
    from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv
    from alphatank_b200.configs import config_teams
    env = MultiAgentEnv(mode="bot_agent", type="train", bot_type="smart", weakness=1.0, num_envs=8, seed=1)
    obs, info = env.reset()
    assert tuple(obs.shape) == (8, 2, 388) and env.observation_space.shape[0] == 776          # F2, H11
    assert list(env.action_space.nvec[:3]) == [3, 3, 2] and env.num_tanks == 2
    acts = torch.zeros((8, 2, 3), dtype=torch.int64, device="cuda")                              # Categorical samples are int64
    obs, rew, done, trunc, info = env.step(acts)
    assert tuple(obs.shape) == (8, 776) and tuple(rew.shape) == (8, 2) and trunc is False
    assert set(info.keys()) == {"change_time", "winner"}
    env.render(); env.set_bot_type("aggressive")
    assert env.game_env.bot_type == "aggressive" and env.game_env.tanks[0].alive.shape == (8,)
    tenv = MultiAgentTeamEnv(config_teams.team_vs_bot_configs, num_envs=4)
    o, _ = tenv.reset()
    assert tuple(o.shape) == (4, 1056) and tenv.num_agents == 2 and tenv.get_observation_order() == ["Tank1", "Tank2"]
    o, r, d, _, info = tenv.step(np.zeros((4, 2, 3), np.int64))
    assert tuple(info["hits"].shape) == (4, 4)
    with pytest.raises(ValueError):
        MultiAgentEnv(mode="human_play", num_envs=2)
    with pytest.raises(ValueError):
        MultiAgentEnv(mode="bot_agent", bot_type="nonsense", num_envs=2)                          # bot_factory.py:29-31
    with pytest.raises(ValueError):
        MultiAgentTeamEnv(config_teams.mixed_team_configs, num_envs=2)                            # human tanks


def test_set_weakness_hot_swap_equals_construction_and_oracle():
    """train_ppo_cycle.py's curriculum changes the opponent's act probability while training: at_set_weakness on a
    live handle gives the stream of a handle constructed with that value (and of the oracle)."""
    from alphatank_b200 import _capi
    spec = dict(SWEEP["botagent_dodge_weak"])
    N, seed = 300, 17
    a = _gpu(dict(spec, weakness=0.35), N, seed, 0, auto_reset=True)
    b = _gpu(dict(spec, weakness=1.0), N, seed, 0, auto_reset=True)
    _capi.check(b.core.L.at_set_weakness(b.core.h, 0.35), "at_set_weakness")
    ora = parity.make_oracle_env(dict(spec, weakness=0.35), N, seed, 0, auto_reset=True)
    r0 = a.reset()
    assert np.array_equal(r0, b.reset()) and np.array_equal(r0, ora.reset())
    ids = np.arange(N)
    for t in range(60):
        acts = synthetic_actions(seed, ids, t, a.A)
        oa, ra, da = a.step(acts)
        ob, rb, db = b.step(acts)
        oo, ro, do = ora.step(acts, threads=4)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db), "tick %d" % t
        assert np.array_equal(oa, oo) and np.array_equal(ra, ro) and np.array_equal(da, do), "oracle, tick %d" % t
    with pytest.raises(ValueError):
        _capi.check(b.core.L.at_set_weakness(b.core.h, 1.5), "at_set_weakness")


@pytest.mark.parametrize("name,knobs", [
    ("team_vs_bot", {}), ("team_vs_bot", {"AT_ROWS_LGG": "0"}), ("team_vs_bot", {"AT_ROWS_LGG": "2", "AT_NO_PDL": "1"}),
    ("team_vs_bot", {"AT_ROWS_BPS": "1", "AT_ROWS_WARPS": "7"}), ("agent_t96", {}), ("agent_t96", {"AT_ROWS_LGG": "3"}),
    ("team_vs_bot_maze", {"AT_ROWS_LGG": "2"}),
    ("team_8_mixed", {}), ("team_three_teams_nodebuff", {}), ("botagent_smart_maze", {}), ("agent_nobuff_battlefield", {}),
])
def test_rows_kernel_launch_shapes_agree_with_the_oracle(name, knobs):
    """rows_kernel (rows rebuilt from the stored state, one bulk copy per group of envs) under its launch knobs -
    envs per group, plain instead of programmatic dependent launch, blocks per SM, warps per block - against the
    oracle: the same bits in every observation column on a ragged batch (1003 envs: the last group is partial),
    through auto-resets, and after a masked reset."""
    spec = SWEEP[name]
    N, seed, base = 1003, 31, 900
    old = {k: os.environ.get(k) for k in ("AT_ROWS_LGG", "AT_NO_PDL", "AT_ROWS_BPS", "AT_ROWS_WARPS")}
    try:
        os.environ.update(knobs)
        b = _gpu(spec, N, seed, base, auto_reset=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
    assert np.array_equal(a.reset(), b.reset())
    ids = np.arange(base, base + N)
    for t in range(120):
        acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
        oa, ra, da = a.step(acts, threads=4)
        ob, rb, db = b.step(acts)
        assert np.array_equal(da, db) and np.array_equal(ra, rb)
        if not np.array_equal(oa, ob):
            raise AssertionError("%s tick %d obs differ at (env, col) %s" % (name, t, np.argwhere(oa != ob)[:8].tolist()))
    mask = (np.arange(N) % 5 == 0).astype(np.uint8)
    mask[-1] = 1
    assert np.array_equal(a.reset(mask), b.reset(mask))


@pytest.mark.parametrize("name", ["team_vs_bot", "agent_t96", "team_vs_bot_maze", "team_8_mixed", "team_harder_battlefield"])
def test_step_norm_equals_step_then_tick_normalize(name):
    """at_step_norm (rows tick-normalised by rows_kernel while still in shared memory) against at_step followed by
    at_tick_normalize: bit-identical, with and without the raw rows, on a ragged batch and through auto-resets."""
    import torch
    from alphatank_b200.learner.ppo_utils import tick_normalize_cuda
    spec = SWEEP[name]
    N, seed, base = 1003, 12, 40
    a, b, c, d, e2 = (_gpu(spec, N, seed, base, auto_reset=True) for _ in range(5))
    for e in (a, b, c, d, e2):
        e.reset()
    ids = np.arange(base, base + N)
    R, D = a.R, a.D
    norm_b = torch.full((N, R * D), -7.0, device="cuda:0")
    norm_c = torch.full((N, R * D), -7.0, device="cuda:0")
    raw_b = torch.full((N, R * D), -7.0, device="cuda:0")
    half_d = torch.full((N, R * D), -7.0, device="cuda:0", dtype=torch.float16)     # at_step_norm16: fp16 rows ...
    half_e = torch.full((N, R * D), -7.0, device="cuda:0", dtype=torch.float16)     # ... with the raw rows beside them
    raw_e = torch.full((N, R * D), -7.0, device="cuda:0")
    for t in range(100):
        acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
        ta = None if acts is None else torch.as_tensor(np.ascontiguousarray(acts, dtype=np.int32))
        oa, ra, da = a.core.step(ta)
        want = tick_normalize_cuda(oa.view(N, R, D)).view(N, R * D)
        _, rb, db = b.core.step_norm(ta, norm_b, obs_out=raw_b)
        _, rc, dc = c.core.step_norm(ta, norm_c)
        assert torch.equal(raw_b, oa) and torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(ra, rc)
        for got in (norm_b, norm_c):
            if not torch.equal(got, want):
                bad = (got != want).nonzero()[:6].tolist()
                raise AssertionError("%s tick %d normalised rows differ at (env, col) %s" % (name, t, bad))
        _, rd, dd = d.core.step_norm16(ta, half_d)
        _, re_, de = e2.core.step_norm16(ta, half_e, obs_out=raw_e)
        assert torch.equal(raw_e, oa) and torch.equal(ra, rd) and torch.equal(da, dd) and torch.equal(ra, re_) and torch.equal(da, de)
        for got in (half_d, half_e):     # the same fp32 values, rounded to nearest fp16
            if not torch.equal(got, want.half()):
                bad = (got != want.half()).nonzero()[:6].tolist()
                raise AssertionError("%s tick %d fp16 normalised rows differ at (env, col) %s" % (name, t, bad))


@pytest.mark.parametrize("name", ["team_vs_bot", "team_harder_battlefield", "agent_t96", "team_8_mixed", "agent_maze"])
def test_persistent_buffer_mode_writes_the_same_rows(name):
    """at_set_rows_mode(1): the step calls ship only the dynamic runs of every row (two bulk copies per row instead of one
    per group) into buffers that were written in full once - raw fp32 rows, normalised fp32 rows (second image and in
    place) and normalised fp16 rows must equal the whole-row writes of handles without the switch, on a ragged batch,
    through auto-resets, with two alternating output buffers.  (Per-env mazes ignore the switch.)"""
    import torch
    spec = SWEEP[name]
    N, seed, base = 1003, 21, 7
    envs = [[_gpu(spec, N, seed, base, auto_reset=True) for _ in range(3)] for _ in range(2)]   # [full | persistent][kind]
    for e in envs[0] + envs[1]:
        e.reset()
    ids = np.arange(base, base + N)
    R, D, A = envs[0][0].R, envs[0][0].D, envs[0][0].A
    dev = "cuda:0"
    mk = lambda dt=torch.float32: [[torch.full((N, R * D), -7.0, device=dev, dtype=dt) for _ in range(2)] for _ in range(2)]
    raw, n2, ni, h = mk(), mk(), mk(), mk(torch.float16)
    for t in range(90):
        acts = synthetic_actions(seed, ids, t, A) if A else None
        ta = None if acts is None else torch.as_tensor(np.ascontiguousarray(acts, dtype=np.int32))
        b = t & 1
        if t == 2:                      # both buffers of every kind have been written in full once
            for e in envs[1]:
                e.core.set_rows_mode(True)
        for m in range(2):
            envs[m][0].core.step_norm(ta, n2[m][b], obs_out=raw[m][b])
            envs[m][1].core.step_norm(ta, ni[m][b])
            envs[m][2].core.step_norm16(ta, h[m][b])
        for x, what in ((raw, "raw"), (n2, "norm (second image)"), (ni, "norm (in place)"), (h, "fp16")):
            if not torch.equal(x[0][b], x[1][b]):
                bad = (x[0][b] != x[1][b]).nonzero()[:6].tolist()
                raise AssertionError("%s tick %d: %s rows differ at (env, col) %s" % (name, t, what, bad))
    # the switch really skips the static block: a buffer that was never written keeps its filler there (fixed maps only)
    lo, hi = envs[1][0].core.static_obs_range()
    fresh = torch.full((N, R * D), -7.0, device=dev)
    ta = None if not A else torch.as_tensor(np.ascontiguousarray(synthetic_actions(seed, ids, 90, A), dtype=np.int32))
    envs[1][0].core.step(ta, obs_out=fresh)
    rows = fresh.view(N, R, D)
    if hi > lo and D % 4 == 0:
        a4, b4 = (lo + 3) & ~3, hi & ~3
        whole = (N // 8) * 8              # (the ragged last group takes the plain-store path and writes whole rows)
        assert (rows[:whole, :, a4:b4] == -7.0).all() and (rows[:, :, :lo] != -7.0).any() and (rows[:, :, b4:] != -7.0).any()
    else:
        assert (rows != -7.0).all() or name == "agent_maze"


@pytest.mark.parametrize("cap", ["1", "2", "3"])
def test_simulation_specialisations_equal_the_generic_kernel(cap):
    """env_kernel<true, false, SPEC>: the simulation kernel compiled for a config class (SimT<SPEC>: team mode + smart
    bots + fixed map; + four tanks; + the 2a_vs_2b layout) against lower levels selected with AT_NO_SPEC (1: the
    generic kernel, 2: SPEC 1, 3: SPEC 2) - same state, rewards, done flags and rows, through auto-resets."""
    spec = SWEEP["team_vs_bot"]
    N, seed, base = 1500, 17, 64
    old = os.environ.get("AT_NO_SPEC")
    try:
        os.environ.pop("AT_NO_SPEC", None)
        a = _gpu(spec, N, seed, base, auto_reset=True)           # SPEC 3
        os.environ["AT_NO_SPEC"] = cap
        b = _gpu(spec, N, seed, base, auto_reset=True)
    finally:
        if old is None:
            os.environ.pop("AT_NO_SPEC", None)
        else:
            os.environ["AT_NO_SPEC"] = old
    assert np.array_equal(a.reset(), b.reset())
    ids = np.arange(base, base + N)
    for t in range(300):
        acts = synthetic_actions(seed, ids, t, a.A)
        oa, ra, da = a.step(acts)
        ob, rb, db = b.step(acts)
        assert np.array_equal(da, db) and np.array_equal(ra, rb) and np.array_equal(oa, ob), "tick %d" % t
    sa, sb = a.get_state(), b.get_state()
    for f in ("n_bullets", "tick", "training_step", "episode_steps", "rng_ctr", "zones"):
        assert np.array_equal(sa[f], sb[f]), f
    for f in ("x", "y", "reward", "angle", "alive", "num_hit", "num_be_hit"):
        assert np.array_equal(sa["tanks"][f][:, :4], sb["tanks"][f][:, :4]), f


@pytest.mark.parametrize("name", ["agent_t96", "agent_nobuff_battlefield", "botagent_smart_weak", "botagent_smart_battlefield_t128"])
def test_1v1_specialisations_equal_the_generic_kernel(name):
    """SimT<4> (1v1 agent mode) and SimT<5> (1v1 bot_agent against the smart bot) on fixed maps against the generic
    kernel (AT_NO_SPEC=1): same rewards, done flags, rows and state through auto-resets and TERMINATE_TIME."""
    spec = SWEEP[name]
    N, seed, base = 1500, 23, 64
    old = os.environ.get("AT_NO_SPEC")
    try:
        os.environ.pop("AT_NO_SPEC", None)
        a = _gpu(spec, N, seed, base, auto_reset=True)
        os.environ["AT_NO_SPEC"] = "1"
        b = _gpu(spec, N, seed, base, auto_reset=True)
    finally:
        if old is None:
            os.environ.pop("AT_NO_SPEC", None)
        else:
            os.environ["AT_NO_SPEC"] = old
    assert np.array_equal(a.reset(), b.reset())
    ids = np.arange(base, base + N)
    for t in range(300):
        acts = synthetic_actions(seed, ids, t, a.A)
        oa, ra, da = a.step(acts)
        ob, rb, db = b.step(acts)
        assert np.array_equal(da, db) and np.array_equal(ra, rb) and np.array_equal(oa, ob), "tick %d" % t
    sa, sb = a.get_state(), b.get_state()
    for f in ("n_bullets", "tick", "training_step", "episode_steps", "rng_ctr", "zones"):
        assert np.array_equal(sa[f], sb[f]), f
    for f in ("x", "y", "reward", "angle", "alive", "num_hit", "num_be_hit"):
        assert np.array_equal(sa["tanks"][f][:, :2], sb["tanks"][f][:, :2]), f


def test_delta_reward_mode_telescopes_to_the_reference_cumulative_rewards():
    """reward_mode="delta" (both reward forms): the per-step increments of the reference's cumulative
    per-episode rewards (quirk F1) - summed over an episode they give the cumulative stream back; explicit masked
    resets restart the sum."""
    import torch
    from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv
    from alphatank_b200.configs.config_teams import team_vs_bot_configs
    dev = torch.device("cuda:0")
    N = 2048
    for make, ar in ((lambda **kw: MultiAgentTeamEnv(team_vs_bot_configs, num_envs=N, device=dev, seed=9, **kw), True),
                     (lambda **kw: MultiAgentEnv(mode="bot_agent", bot_type="aggressive", num_envs=N, device=dev, seed=9, **kw), False)):
        a = make(auto_reset=ar)
        b = make(auto_reset=ar, reward_mode="delta")
        a.reset(); b.reset()
        acc = torch.zeros((N, a.core.R), dtype=torch.float64, device=dev)
        ended = 0
        for t in range(300):
            acts = a.core.synthetic_actions(t)
            _, ra, da, _, _ = a.step(acts)
            _, rb, db, _, _ = b.step(acts)
            assert torch.equal(da, db)
            acc += rb.double()
            assert torch.allclose(acc, ra.double(), rtol=0, atol=5e-3), t
            d = da.bool()
            ended += int(d.sum())
            acc[d] = 0.0
            if not ar:                       # the reference's protocol: the caller resets finished envs
                a.reset(mask=da); b.reset(mask=db)
        assert ended > 50
        with pytest.raises(ValueError):
            make(reward_mode="per_tick")
        a.close(); b.close()


def test_kernel_delta_rewards_equal_the_difference_of_consecutive_cumulative_rewards():
    """AT_REWARD_DELTA is computed inside the step (csrc/at_sim.h finish_step) from the previous value kept per env: the
    same float32 numbers as subtracting consecutive cumulative outputs on the caller's side, zero-based after a reset."""
    import torch
    from alphatank_b200 import MultiAgentTeamEnv
    from alphatank_b200.configs.config_teams import team_vs_bot_configs
    dev = torch.device("cuda:0")
    N = 1500
    a = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=N, device=dev, seed=4)
    b = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=N, device=dev, seed=4, reward_mode="delta")
    a.reset(); b.reset()
    prev = torch.zeros((N, 2), dtype=torch.float32, device=dev)
    for t in range(260):
        acts = a.core.synthetic_actions(t)
        _, ra, da, _, _ = a.step(acts)
        _, rb, db, _, _ = b.step(acts)
        assert torch.equal(rb, ra - prev), t
        prev = ra * (1.0 - da.to(torch.float32)).unsqueeze(-1)
    b.core.set_reward_mode("cumulative")          # and back: the same stream as the cumulative handle again
    acts = a.core.synthetic_actions(260)
    _, ra, _, _, _ = a.step(acts)
    _, rb, _, _, _ = b.step(acts)
    assert torch.equal(ra, rb)


def test_observe_and_tank_columns_read_the_current_state():
    """at_observe = env._get_observation() on the current state (the rows the last step returned); the TankView
    properties (env.game_env.tanks[i].alive ..., inference.py:115-116) come from one column kernel and agree with the
    full-state snapshot."""
    import torch
    from alphatank_b200 import MultiAgentEnv
    env = MultiAgentEnv(mode="bot_agent", bot_type="smart", num_envs=777, device="cuda:0", seed=2)
    env.reset()
    for t in range(90):
        obs, _, _, _, _ = env.step(env.core.synthetic_actions(t))
    last = obs.clone()
    again = env.core.observe(torch.empty_like(env.core.obs))
    assert torch.equal(last.view(-1), again.view(-1))
    st = env.core.get_state()
    for i in range(2):
        tv = env.game_env.tanks[i]
        assert np.array_equal(tv.alive.cpu().numpy(), st["tanks"]["alive"][:, i] != 0)
        assert np.array_equal(tv.x.cpu().numpy(), st["tanks"]["x"][:, i])
        assert np.array_equal(tv.angle.cpu().numpy(), st["tanks"]["angle"][:, i])
        assert np.array_equal(tv.reward.cpu().numpy(), st["tanks"]["reward"][:, i])
        assert np.array_equal(tv.num_be_hit.cpu().numpy(), st["tanks"]["num_be_hit"][:, i])
    assert tv.alive.is_cuda


def test_bot_agent_outside_training_reads_the_agent_action_from_row_0():
    """gaming_env.py:176-177 (F15): MultiAgentEnv(mode='bot_agent', type != 'train') - the inference scripts - takes the
    agent's action from actions[0].  Same trajectories as a training-type env fed the same action in row 1."""
    import torch
    from alphatank_b200 import MultiAgentEnv
    N = 600
    a = MultiAgentEnv(mode="bot_agent", type="train", bot_type="aggressive", num_envs=N, device="cuda:0", seed=8)
    b = MultiAgentEnv(mode="bot_agent", type="inference", bot_type="aggressive", num_envs=N, device="cuda:0", seed=8)
    a.reset(); b.reset()
    for t in range(150):
        acts = a.core.synthetic_actions(t).clone()
        oa, ra, da, _, _ = a.step(acts)
        ob, rb, db, _, _ = b.step(acts[:, 1:2])          # one row per env, as inference.py passes it
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t


def test_handles_with_different_shared_memory_needs_coexist():
    """The dynamic shared-memory limit of a kernel is process-wide: a second handle with a smaller working set
    (octagon: 40 walls) must not lower it under a first one that needs more (battlefield: 61 walls)."""
    big = _gpu(SWEEP["team_harder_battlefield"], 300, 5, 0, auto_reset=True)
    small = _gpu(SWEEP["team_vs_bot"], 300, 5, 0, auto_reset=True)
    ora = parity.make_oracle_env(SWEEP["team_harder_battlefield"], 300, 5, 0, auto_reset=True)
    assert np.array_equal(big.reset(), ora.reset())
    small.reset()
    ids = np.arange(300)
    for t in range(20):
        acts = synthetic_actions(5, ids, t, big.A)
        ob, rb, db = big.step(acts)
        small.step(synthetic_actions(5, ids, t, small.A))
        oo, ro, do = ora.step(acts, threads=4)
        assert np.array_equal(ob, oo) and np.array_equal(rb, ro) and np.array_equal(db, do), t


def test_set_weakness_reaches_a_captured_cuda_graph():
    """at_set_weakness writes device memory the kernels read (DevState::dyn), so a step captured in a CUDA graph - the
    rollout runner's case - follows the curriculum: after set_weakness(0) the replayed graph's bot never acts again."""
    import torch
    from alphatank_b200 import MultiAgentEnv
    N = 512
    env = MultiAgentEnv(mode="bot_agent", bot_type="aggressive", weakness=1.0, num_envs=N, device="cuda:0", seed=6,
                        auto_reset=False)
    env.reset()
    acts = torch.ones((N, 2, 3), dtype=torch.int32, device="cuda:0")     # the agent idles
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            env.core.step(acts)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            env.core.step(acts)
    torch.cuda.synchronize()
    for _ in range(40):
        g.replay()
    torch.cuda.synchronize()
    moved = env.core.get_state()["tanks"]
    assert (moved["speed"][:, 0] != 0).any() or (moved["num_hit"][:, 0] > 0).any()   # weakness 1: the bot acts
    env.set_weakness(0.0)
    before = env.core.get_state()["tanks"]
    for _ in range(60):
        g.replay()
    torch.cuda.synchronize()
    after = env.core.get_state()["tanks"]
    assert np.array_equal(before["x"][:, 0], after["x"][:, 0]) and np.array_equal(before["angle"][:, 0], after["angle"][:, 0])
    assert (after["speed"][:, 0] == 0).all()
