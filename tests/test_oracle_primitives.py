"""CPU: known-answer tests of the oracle's primitives against golden vectors recorded from the
reference (tests/golden/bfs.npz, mazes.npz) and the anchors of SURVEY.md App. H."""
import math
import os

import numpy as np

import parity
from oracle import oracle as O
from oracle.atrng import philox4x32, word


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    def h(*a):
        return tuple(int(x) for x in philox4x32(*a))
    assert h(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert h(f, f, f, f, f, f) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert h(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_c_rng_matches_numpy_rng():
    L = O.lib()
    for seed, env, stream, ctr in [(0, 0, 0, 0), (7, 3, 0, 11), (2**40 + 5, 2**33 + 1, 2, 2**31), (99, 65535, 1, 77)]:
        assert L.ato_rng_word(seed, env, stream, ctr) == int(word(seed, env, stream, ctr)) or stream != 0
        w = philox4x32(ctr, stream, env & 0xffffffff, env >> 32, seed & 0xffffffff, seed >> 32)[0]
        assert L.ato_rng_word(seed, env, stream, ctr) == int(w)


def test_bfs_matches_reference_on_all_pairs():
    z = np.load(os.path.join(parity.GOLDEN, "bfs.npz"))
    grids, cases = z["grids"], z["cases"]
    L = O.lib()
    out = np.zeros(2 * O.CELLS, np.int32)
    flat = [np.ascontiguousarray(g.reshape(-1)) for g in grids]
    for gi, sr, sc, gr, gc, ln, nr, nc in cases.tolist():
        n = L.ato_bfs_path(flat[gi].ctypes.data, sr, sc, gr, gc, out.ctypes.data)
        assert n == ln
        if ln > 1:
            assert (out[2], out[3]) == (nr, nc)


def test_bfs_anchor_h1():
    oct_ = np.ones((11, 11), np.uint8)
    oct_[1:-1, 1:-1] = 0
    assert O.bfs_path(oct_, (1, 1), (3, 4)) == [(1, 1), (2, 1), (3, 1), (3, 2), (3, 3), (3, 4)]
    assert O.bfs_path(oct_, (1, 1), (1, 1)) == [(1, 1)]
    assert O.bfs_path(oct_, (0, 0), (1, 1)) is None       # start in a wall
    assert O.bfs_path(oct_, (1, 1), (11, 1)) is None      # goal off the map


def test_mazes_match_reference():
    z = np.load(os.path.join(parity.GOLDEN, "mazes.npz"))
    for i, (m, c) in enumerate(zip(z["mazes"], z["ctrs"])):
        got, ctr = O.generate_maze(int(z["seed"]), i)
        assert np.array_equal(got, m) and ctr == c
        assert int((got == 0).sum()) == 49 and got[0].all() and got[-1].all() and got[:, 0].all() and got[:, -1].all()


def test_trig_is_python_math():
    L = O.lib()
    import ctypes as C
    c, s = C.c_double(), C.c_double()
    for a in range(361):
        L.ato_trig(a, C.byref(c), C.byref(s))
        assert c.value == math.cos(math.radians(a)) and s.value == math.sin(math.radians(a))


def _one_env(mode="agent", terminate_time=None):
    spec = dict(mode=mode, teams=[0, 1], ctrl=["agent", "agent"], bots=[None, None], map="octagon",
                terminate_time=terminate_time, buff_on=True, debuff_on=True, weakness=1.0)
    e = parity.make_oracle_env(spec, 1, 1)
    st = e.get_state(0).copy()
    st["zones"][:] = [70, 70, 70, 70, 70, 70, 70, 70]
    return e, st


def _place(st, i, x, y, angle):
    t = st["tanks"][i]
    t["x"], t["y"], t["angle"], t["speed"], t["alive"], t["reward"] = x, y, angle, 0, 1, 0.0


NOOP, FWD, SHOOT = [1, 1, 0], [1, 2, 0], [1, 1, 1]


def test_anchor_h2_wall_slack_and_h3_move():
    e, st = _one_env()
    _place(st, 0, 105.0, 105.0, 180); _place(st, 1, 385.0, 385.0, 30)
    st["tick"] = 100
    e.set_state(0, st)
    xs = []
    for _ in range(4):
        e.step([[FWD, FWD if not xs else NOOP]])
        s = e.get_state(0)
        xs.append((float(s["tanks"]["x"][0]), int(s["tanks"]["hitting_wall"][0])))
    assert xs == [(95.0, 0), (95.0, 1), (95.0, 1), (95.0, 1)]
    assert float(s["tanks"]["y"][0]) == 105.0
    assert (float(s["tanks"]["x"][1]), float(s["tanks"]["y"][1])) == (393.66025403784437, 380.0)


def test_anchor_h4_sat_epsilon():
    L = O.lib()
    right = (700, 350, 70, 70)
    assert L.ato_obb_vs_aabb(684.98, 385.0, 0, *right) == 0
    assert L.ato_obb_vs_aabb(684.995, 385.0, 0, *right) == 1
    assert L.ato_obb_vs_aabb(685.0, 385.0, 0, *right) == 1


def test_anchor_h5_cooldown_on_virtual_clock():
    # ticks 17 / 18 -> 283 / 300 ms on the virtual clock: no shot / shot (sprite.py:593)
    e, st = _one_env()
    _place(st, 0, 385.0, 385.0, 30); _place(st, 1, 105.0, 105.0, 0)
    st["tick"] = 16
    e.set_state(0, st)
    e.step([[SHOOT, NOOP]])
    assert e.get_state(0)["n_bullets"] == 0
    e.step([[SHOOT, NOOP]])
    s = e.get_state(0)
    assert s["n_bullets"] == 1
    b = s["bullets"][0]
    # H3: spawn point + direction, then one pass-2 sub-step
    assert (float(b["dx"]), float(b["dy"])) == (0.8660254037844387, -0.49999999999999994)
    assert float(b["x"]) == 393.66025403784437 + 0.8660254037844387 and float(b["y"]) == 380.0 + -0.49999999999999994


def _fly(e, n_ticks):
    log = []
    for _ in range(n_ticks):
        e.step([[NOOP, NOOP]])
        s = e.get_state(0)
        log.append(s.copy())
        if s["n_bullets"] == 0:
            break
    return log


def _shoot_from(x, y, angle, ex, ey, terminate_time=None):
    e, st = _one_env(terminate_time=terminate_time)
    _place(st, 0, x, y, angle); _place(st, 1, ex, ey, 0)
    st["tick"] = 1000
    e.set_state(0, st)
    e.step([[SHOOT, NOOP]])
    return e


def test_anchor_h6_two_bounces_then_expiry():
    e = _shoot_from(105.0, 105.0, 100, 665.0, 665.0)
    log = _fly(e, 200)
    seen = {}
    for s in [e.get_state(0)] + log:
        pass
    # replay again, recording the sub-step index of each bounce through distance/bounces
    e = _shoot_from(105.0, 105.0, 100, 665.0, 665.0)
    prev_b, sub = 0, 1
    while True:
        s = e.get_state(0)
        if s["n_bullets"] == 0:
            break
        b = s["bullets"][0]
        if b["bounces"] != prev_b:
            seen[int(b["bounces"])] = int(b["distance"])
            prev_b = int(b["bounces"])
        e.step([[NOOP, NOOP]])
    # bounce k happened within the 2 sub-steps of the tick where it was first observed
    assert 26 <= seen[1] <= 27 and 192 <= seen[2] <= 193
    s = e.get_state(0)
    assert tuple(s["tanks"]["reward"][:2]) == (0.0, 0.0)


def test_anchor_h7_corner_trap():
    e = _shoot_from(105.0, 105.0, 135, 665.0, 665.0)
    last = None
    for _ in range(60):
        s = e.get_state(0)
        if s["n_bullets"] == 0:
            break
        last = s["bullets"][0].copy()
        e.step([[NOOP, NOOP]])
    assert e.get_state(0)["n_bullets"] == 0
    assert int(last["distance"]) <= 45 and int(last["bounces"]) >= 4   # removed at sub-step 45 with 6 bounces


def test_anchor_h8_hit_rewards_and_death():
    e = _shoot_from(105.0, 385.0, 0, 245.0, 385.0)
    for _ in range(80):
        e.step([[NOOP, NOOP]])
        if e.done[0]:
            break
    s = e.get_state(0)
    assert e.done[0] == 1 and s["tanks"]["alive"][1] == 0 and s["tanks"]["num_hit"][0] == 1
    assert abs(float(s["tanks"]["reward"][0]) - 51.11) < 1e-9 and float(s["tanks"]["reward"][1]) == -50.0
    assert e.info()["winner"][0] == 0


def test_anchor_h9_away_shot_penalty_cancelled():
    e = _shoot_from(105.0, 385.0, 180, 385.0, 385.0)
    for _ in range(200):
        e.step([[NOOP, NOOP]])
        if e.get_state(0)["n_bullets"] == 0:
            break
    s = e.get_state(0)
    assert abs(float(s["tanks"]["reward"][0]) - 2.58) < 1e-9 and s["tanks"]["alive"][1] == 1


def test_f6_victory_bonus_on_every_hit_when_immortal():
    e = _shoot_from(105.0, 385.0, 0, 245.0, 385.0, terminate_time=4096)
    for _ in range(80):
        e.step([[NOOP, NOOP]])
        if e.get_state(0)["n_bullets"] == 0:
            break
    s = e.get_state(0)
    steps = int(s["episode_steps"])
    assert s["tanks"]["alive"][1] == 1 and s["tanks"]["num_be_hit"][1] == 1
    bonus = 200 * (5 * ((1000 - steps) / 1000))          # gaming_env.py:535-538, paid to BOTH tanks
    assert float(s["tanks"]["reward"][1]) == -50 + bonus
    assert abs(float(s["tanks"]["reward"][0]) - (51.11 + bonus)) < 1e-9
    assert steps < 80
