"""CPU: the CUDA kernel's per-thread code (alphatank_b200/csrc/at_sim.h), compiled with g++ into the
test-only emulator (tests/cpu_emul), against the reference traces and against the oracle.

This checks the kernel FORMULATION (bitmask walls, LUT trig, packed bullets, bit-parallel BFS, row builders)
in the GPU-less build container; the real kernels are checked by the -m gpu tests."""
import os

import numpy as np
import pytest

import emul_env
import parity
from oracle import oracle as O
from oracle.atrng import synthetic_actions


@pytest.mark.parametrize("name", parity.fixture_names())
def test_kernel_logic_replays_reference_trace(name):
    rec = parity.load_fixture(name)
    meta = rec["meta"]
    env = emul_env.make_emul_env(parity.env_spec(meta), 1, meta["seed"], meta["env_id"])
    assert env.D * env.R == rec["obs"].shape[1]
    parity.replay_fixture(env, rec, where=name, max_ticks=700, check_state_every=3)


def _spec(mode, layout, map_="octagon", tt=None, weakness=1.0, buff=True, debuff=True):
    ids = {}
    return dict(mode=mode, teams=[ids.setdefault(t, len(ids)) for t, _, _ in layout], ctrl=[c for _, c, _ in layout],
                bots=[b for _, _, b in layout], map=map_, terminate_time=tt, buff_on=buff, debuff_on=debuff,
                weakness=weakness)


T = parity.TEAM_DICTS
SWEEP = {
    "agent_maze": _spec("agent", [("A", "agent", None), ("B", "agent", None)], "maze"),
    "agent_nobuff_battlefield": _spec("agent", [("A", "agent", None), ("B", "agent", None)], "battlefield", buff=False),
    "agent_t96": _spec("agent", [("A", "agent", None), ("B", "agent", None)], tt=96),
    "botagent_smart_maze": _spec("bot_agent", [("A", "bot", "smart"), ("B", "agent", None)], "maze"),
    "botagent_dodge_weak": _spec("bot_agent", [("A", "bot", "dodge"), ("B", "agent", None)], weakness=0.6),
    "botagent_smart_weak": _spec("bot_agent", [("A", "bot", "smart"), ("B", "agent", None)], weakness=0.8),            # SPEC 5
    "botagent_smart_battlefield_t128": _spec("bot_agent", [("A", "bot", "smart"), ("B", "agent", None)], "battlefield", tt=128),
    "arena_def_dodge_t200": _spec("arena", [("A", "bot", "defensive"), ("B", "bot", "dodge")], tt=200),
    "arena_random_aggr_maze": _spec("arena", [("A", "bot", "random"), ("B", "bot", "aggressive")], "maze"),
    "team_vs_bot": _spec("team", T["team_vs_bot_configs"]),
    "team_vs_bot_maze": _spec("team", T["team_vs_bot_configs"], "maze"),
    "team_harder_battlefield": _spec("team", T["team_vs_bot_harder_configs"], "battlefield"),
    "team_three_teams_nodebuff": _spec("team", T["bot_team_configs"], debuff=False),
    "team_8_mixed": _spec("team", [("A", "agent", None), ("A", "bot", "dodge"), ("B", "bot", "smart"), ("B", "agent", None),
                                   ("C", "bot", "random"), ("C", "bot", "defensive"), ("D", "bot", "aggressive"),
                                   ("D", "agent", None)], tt=150),
}


@pytest.mark.parametrize("name", sorted(SWEEP))
def test_kernel_logic_matches_oracle_on_batches(name):
    spec = SWEEP[name]
    N, ticks, seed, base = 96, 260, 2024, 1000   # 96 envs: one full and one partial block of 64
    a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
    b = emul_env.make_emul_env(spec, N, seed, base, auto_reset=True)
    assert (a.D, a.R, a.A) == (b.D, b.R, b.A)
    assert np.array_equal(a.reset(), b.reset())
    ids = np.arange(base, base + N)
    dones = 0
    for t in range(ticks):
        acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
        oa, ra, da = a.step(acts, threads=4)
        ob, rb, db = b.step(acts)
        assert np.array_equal(da, db), "%s tick %d done" % (name, t)
        assert np.array_equal(ra, rb), "%s tick %d rewards" % (name, t)
        if not np.array_equal(oa, ob):
            bad = np.argwhere(oa != ob)
            raise AssertionError("%s tick %d obs differ at (env, col) %s" % (name, t, bad[:8].tolist()))
        if da.any() or t == 0:   # info of the step as the reference's callers see it: the finished episode's, latched before the auto-reset
            ta, tb = a.terminal_info(), b.terminal_info()
            for f in ta:
                assert np.array_equal(ta[f], tb[f]), "%s tick %d terminal info %s" % (name, t, f)
            assert np.array_equal(tb["latched"], db)
        dones += int(da.sum())
    sa, sb = a.get_state(), b.get_state()
    for f in ("n_bullets", "tick", "training_step", "episode_steps", "rng_ctr", "zones", "maze"):
        assert np.array_equal(sa[f], sb[f]), f
    n = a.n_tanks
    for f in ("x", "y", "reward", "angle", "speed", "alive", "hitting_wall", "in_buff", "in_debuff", "max_bullets",
              "last_shot_ms", "num_hit", "num_be_hit"):
        assert np.array_equal(sa["tanks"][f][:, :n], sb["tanks"][f][:, :n]), f
    for i in range(N):
        nb = sa["n_bullets"][i]
        for f in ("x", "y", "dx", "dy", "pending", "owner", "distance", "bounces", "cleared"):
            assert np.array_equal(sa["bullets"][f][i, :nb], sb["bullets"][f][i, :nb]), f
    ia, ib = a.info(), b.info()
    for f in ia:
        assert np.array_equal(ia[f], ib[f]), f
    assert dones > 0 or spec["mode"] == "agent"


def test_bfs_bitparallel_matches_reference_pairs():
    z = np.load(os.path.join(parity.GOLDEN, "bfs.npz"))
    grids, cases = z["grids"], z["cases"]
    L = emul_env.lib()
    import ctypes as C
    nr, nc = C.c_int(), C.c_int()
    flat = [np.ascontiguousarray(g.reshape(-1)) for g in grids]
    for gi, sr, sc, gr, gc, ln, er, ec in cases[::3].tolist():
        n = L.em_bfs(flat[gi].ctypes.data, sr, sc, gr, gc, C.byref(nr), C.byref(nc))
        assert n == ln
        if ln > 1:
            assert (nr.value, nc.value) == (er, ec)


def test_maze_kernel_logic_matches_reference():
    z = np.load(os.path.join(parity.GOLDEN, "mazes.npz"))
    import ctypes as C
    L = emul_env.lib()
    for i, (m, c) in enumerate(zip(z["mazes"], z["ctrs"])):
        out = np.zeros(121, np.uint8)
        ctr = C.c_uint32(0)
        L.em_generate_maze(int(z["seed"]), i, C.byref(ctr), out.ctypes.data)
        assert np.array_equal(out.reshape(11, 11), m) and ctr.value == c


def test_state_roundtrip_through_packed_layout():
    spec = SWEEP["team_8_mixed"]
    a = parity.make_oracle_env(spec, 4, 5, 0)
    b = emul_env.make_emul_env(spec, 4, 5, 0)
    ids = np.arange(4)
    for t in range(150):
        a.step(synthetic_actions(5, ids, t, a.A))
    for i in range(4):                      # inject oracle state into the packed SoA layout and read it back
        b.set_state(i, a.get_state(i))
        sa, sb = a.get_state(i), b.get_state(i)
        n, nb = a.n_tanks, sa["n_bullets"]
        for f in ("x", "y", "reward", "angle", "speed", "alive", "num_hit", "num_be_hit", "last_shot_ms"):
            assert np.array_equal(sa["tanks"][f][:n], sb["tanks"][f][:n]), f
        for f in ("x", "y", "dx", "dy", "pending", "owner", "distance", "bounces", "cleared"):
            assert np.array_equal(sa["bullets"][f][:nb], sb["bullets"][f][:nb]), f
    for t in range(150, 260):               # and keep stepping in lock-step
        acts = synthetic_actions(5, ids, t, a.A)
        oa, ra, da = a.step(acts)
        ob, rb, db = b.step(acts)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)


def test_exact_multi_step_jump_equals_sequential_additions():
    """jump_axis (one fma inside a binade) must be bit-identical to m sequential `x = x + d` additions - the
    line-of-sight skip-ahead relies on it.  Covers all LUT directions, binade crossings (64, 128, 256, 512),
    tiny / absorbed steps and constructed rounding ties."""
    import math
    L = emul_env.lib()
    rng = np.random.RandomState(0)
    dirs = [s * f(math.radians(a)) for a in range(361) for f in (math.cos, math.sin) for s in (1.0, -1.0)]
    cases = []
    for _ in range(6000):
        cases.append((rng.uniform(60.0, 710.0), dirs[rng.randint(len(dirs))], int(rng.randint(1, 301))))
    for b in (64.0, 128.0, 256.0, 512.0):                      # straddle every binade boundary of the arena
        for _ in range(300):
            cases.append((b + rng.uniform(-40, 40), dirs[rng.randint(len(dirs))], int(rng.randint(4, 120))))
    for e in (6, 7, 8, 9):                                     # exact ties: d = q*u + u/2
        u = 2.0 ** (e - 52)
        for q in (3, 1000, 2 ** 30 + 1):
            for sgn in (1.0, -1.0):
                cases.append((2.0 ** e * 1.37, sgn * (q * u + u / 2), 50))
                cases.append((2.0 ** e * 1.37 + u, sgn * (q * u + u / 2), 50))
    for x, d, m in cases:
        want = x
        for _ in range(m):
            want = want + d
        got = L.em_jump_axis(x, d, m)
        assert got == want, (x, d, m, got, want)


def test_squared_distance_thresholds_are_exact():
    """sqrt(s) > 300 <=> s > S300 and sqrt(s) < 100 <=> s < S100 for every double s (checked on the +-300
    neighbours of each boundary; sqrt is monotonic and correctly rounded)."""
    import math
    import re
    src = open(os.path.join(os.path.dirname(parity.GOLDEN), "..", "alphatank_b200", "csrc", "at_sim.h")).read()
    s300 = float.fromhex(re.search(r"S300 = (0x[0-9a-fp.+]+);", src).group(1))
    s100 = float(re.search(r"S100 = ([0-9.]+);", src).group(1))
    for thr, s0, gt in ((300.0, s300, True), (100.0, s100, False)):
        s = s0
        for _ in range(300):
            s = math.nextafter(s, -math.inf)
        for _ in range(600):
            if gt:
                assert (math.sqrt(s) > thr) == (s > s0), s.hex()
            else:
                assert (math.sqrt(s) < thr) == (s < s0), s.hex()
            s = math.nextafter(s, math.inf)


@pytest.mark.parametrize("cap", [0, 1, 2, 3])
def test_simulation_specialisations_agree(cap):
    """SimT<SPEC> (csrc/at_sim.h): the compile-time specialisations of the simulation kernel (team mode + smart bots +
    fixed map; + four tanks; + the 2a_vs_2b layout) against the oracle on the configs that select them - capped at
    every level, so each specialisation and the generic build run the same 2a_vs_2b batch."""
    specs = [SWEEP["team_vs_bot"],
             _spec("team", [("A", "agent", None), ("B", "bot", "smart"), ("B", "bot", "smart"), ("A", "agent", None)]),   # SPEC 2
             _spec("team", [("A", "agent", None), ("B", "bot", "smart"), ("C", "bot", "smart")], "battlefield"),        # SPEC 1
             SWEEP["agent_t96"], SWEEP["botagent_smart_weak"]]          # SPEC 4 / 5 (any cap below 3: the generic build)
    emul_env.lib().em_set_spec_cap(cap)
    try:
        for spec in specs:
            N, seed, base = 70, 5, 300
            a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
            b = emul_env.make_emul_env(spec, N, seed, base, auto_reset=True)
            assert np.array_equal(a.reset(), b.reset())
            ids = np.arange(base, base + N)
            for t in range(260):
                acts = synthetic_actions(seed, ids, t, a.A)
                oa, ra, da = a.step(acts, threads=4)
                ob, rb, db = b.step(acts)
                assert np.array_equal(da, db) and np.array_equal(ra, rb) and np.array_equal(oa, ob), (cap, t)
    finally:
        emul_env.lib().em_set_spec_cap(3)


@pytest.mark.parametrize("order", [1, 2])
def test_pooled_stages_do_not_depend_on_the_lane_order(order):
    """The pooled line-of-sight stage (Sim::pool_los) runs here on the SIMT emulator (tests/cpu_emul/simt.h): 32
    fibers per warp that meet at the warp collectives.  Between two collectives the emulator runs the lanes one after
    the other - in reverse (1) or freshly shuffled (2) order here instead of lane order - so a lane that read another
    lane's shared-memory write without a barrier in between would diverge from the oracle.  Smart-bot configs: the
    bench layout, a 1v1 arena of two smart bots, the 6-tank battlefield mix."""
    L = emul_env.lib()
    L.em_set_lane_order(order)
    try:
        for name in ("team_vs_bot", "arena_smart_smart", "team_harder_battlefield"):
            spec = SWEEP[name] if name in SWEEP else _spec("arena", [("A", "bot", "smart"), ("B", "bot", "smart")])
            N, ticks, seed, base = 70, 220, 31, 64
            a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
            b = emul_env.make_emul_env(spec, N, seed, base, auto_reset=True)
            assert np.array_equal(a.reset(), b.reset())
            ids = np.arange(base, base + N)
            for t in range(ticks):
                acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
                oa, ra, da = a.step(acts, threads=4)
                ob, rb, db = b.step(acts)
                assert np.array_equal(da, db) and np.array_equal(ra, rb) and np.array_equal(oa, ob), (name, t)
            sa, sb = a.get_state(), b.get_state()
            for f in ("n_bullets", "rng_ctr"):
                assert np.array_equal(sa[f], sb[f]), f
            assert L.em_collectives(b.h) > 0
    finally:
        L.em_set_lane_order(0)


def test_atan2_refine_is_correctly_rounded():
    """atan2_refine (csrc/at_sim.h: one Newton step on a few-ulp atan2, sin / cos by angle addition from a table of
    106-bit values at the integer-degree angles) against mpmath at 200 bits, on the inputs it exists for - bearings that
    are integers up to rounding (a tank that left a shared spawn cell along its heading, seen from that cell) - starting
    from values up to 3 ulp off.  The same run reports how often glibc's atan2 (= Python's = the oracle's) is itself
    correctly rounded there: that rate bounds how often a knife-edge bot decision on the GPU can differ from the
    reference, whose own answer depends on the libm it runs on."""
    mpmath = pytest.importorskip("mpmath")
    import math
    mpmath.mp.prec = 200
    L = emul_env.lib()
    cases = []
    for a in range(0, 361):
        r = math.radians(a)
        c, s = math.cos(r), math.sin(r)
        for k in (1, 2, 3, 5, 8, 13, 21, 29):
            for sp in (10, -10, 1):
                x, y = 105.0, 595.0
                for _ in range(k):
                    x, y = x + sp * c, y - sp * s
                dx, dy = x - 105.0, y - 595.0
                if dx != 0.0 and dy != 0.0 and abs(dx) != abs(dy):   # (axis / diagonal cases are pinned constants)
                    cases += [(-dy, dx), (dy, -dx)]
    bad = bad_libm = refined = 0
    for y, x in cases:
        want = float(mpmath.atan2(mpmath.mpf(y), mpmath.mpf(x)))
        bad_libm += math.atan2(y, x) != want
        for ulps in (0, 2, -3):
            got = L.em_atan2_refine(y, x, ulps)
            if got == got:     # NaN: the bearing is not within 1e-9 degrees of an integer (never refined)
                refined += 1
                bad += got != want
    assert refined > 0.9 * 3 * len(cases)
    assert bad == 0, "%d of %d not correctly rounded" % (bad, refined)
    assert bad_libm < 0.02 * len(cases)      # glibc 2.39: ~0.4 %
