"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; alphatank_b200/ never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

MODE = {"agent": 0, "bot_agent": 1, "arena": 2, "team": 3}
BOT = {"smart": 0, "random": 1, "aggressive": 2, "defensive": 3, "dodge": 4, "stationary": 5}
MAP = {"octagon": 0, "battlefield": 1, "maze": 2}
MAX_TANKS, MAX_BULLETS, CELLS = 8, 48, 121


class Config(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_tanks", C.c_int32), ("map", C.c_int32), ("terminate_time", C.c_int32),
                ("buff_on", C.c_int32), ("debuff_on", C.c_int32), ("auto_reset", C.c_int32),
                ("aim_raymarch", C.c_int32), ("team", C.c_int32 * MAX_TANKS), ("ctrl", C.c_int32 * MAX_TANKS),
                ("bot_type", C.c_int32 * MAX_TANKS), ("weakness", C.c_double)]


TANK_DT = np.dtype([("x", "f8"), ("y", "f8"), ("reward", "f8"), ("angle", "i4"), ("speed", "i4"), ("alive", "i4"),
                    ("hitting_wall", "i4"), ("in_buff", "i4"), ("in_debuff", "i4"), ("max_bullets", "i4"),
                    ("last_shot_ms", "i4"), ("num_hit", "i4"), ("num_be_hit", "i4")], align=True)
BULLET_DT = np.dtype([("x", "f8"), ("y", "f8"), ("dx", "f8"), ("dy", "f8"), ("pending", "f8"), ("owner", "i4"),
                      ("distance", "i4"), ("bounces", "i4"), ("cleared", "i4")], align=True)
BOT_DT = np.dtype([("target", "i4"), ("aiming", "i4"), ("rot_dir", "i4"), ("stuck_timer", "i4"), ("has_next", "i4"),
                   ("next_r", "i4"), ("next_c", "i4"), ("pad0", "i4"), ("last_x", "f8"), ("last_y", "f8"),
                   ("has_tpos", "i4"), ("pad1", "i4"), ("tpos_x", "f8"), ("tpos_y", "f8"), ("dodge_timer", "i4"),
                   ("dodge_state", "i4"), ("has_dodge_angle", "i4"), ("pad2", "i4"), ("dodge_angle", "f8"),
                   ("action_delay", "i4"), ("cur_action", "i4", (3,))], align=True)
ENV_DT = np.dtype([("n_bullets", "i4"), ("tick", "i4"), ("training_step", "i4"), ("episode_steps", "i4"),
                   ("rng_ctr", "u4"), ("n_walls", "i4"), ("zones", "i4", (8,)), ("maze", "u1", (CELLS + 7,)),
                   ("tanks", TANK_DT, (MAX_TANKS,)), ("bots", BOT_DT, (MAX_TANKS,)),
                   ("bullets", BULLET_DT, (MAX_BULLETS,))], align=True)


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("tank_oracle.c", "tank_oracle.h", "Makefile")]
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in src)):
        subprocess.check_call(["make", "-C", HERE, "-s", "liboracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.ato_create.restype = vp
        L.ato_create.argtypes = [C.POINTER(Config), i64, i64, u64]
        L.ato_destroy.argtypes = [vp]
        for f in ("ato_obs_dim", "ato_obs_rows", "ato_action_rows"):
            getattr(L, f).restype = i32
            getattr(L, f).argtypes = [vp]
        L.ato_reset.argtypes = [vp, vp, vp]
        L.ato_step.argtypes = [vp, vp, vp, vp, vp]
        L.ato_step_mt.argtypes = [vp, vp, vp, vp, vp, i32]
        L.ato_get_state.argtypes = [vp, i64, vp]
        L.ato_set_state.argtypes = [vp, i64, vp]
        L.ato_get_info.argtypes = [vp, vp, vp, vp, vp]
        L.ato_get_terminal_info.argtypes = [vp, vp, vp, vp, vp, vp]
        L.ato_bfs_len_between.restype = i32
        L.ato_bfs_len_between.argtypes = [vp, i64, i32, i32]
        L.ato_bfs_path.restype = i32
        L.ato_bfs_path.argtypes = [vp, i32, i32, i32, i32, vp]
        L.ato_generate_maze.argtypes = [u64, i64, vp, vp]
        L.ato_obb_vs_aabb.restype = i32
        L.ato_obb_vs_aabb.argtypes = [C.c_double, C.c_double, i32, i32, i32, i32, i32]
        L.ato_corners.argtypes = [C.c_double, C.c_double, i32, vp]
        L.ato_rng_word.restype = C.c_uint32
        L.ato_rng_word.argtypes = [u64, i64, C.c_uint32, C.c_uint32]
        L.ato_trig.argtypes = [i32, vp, vp]
        assert C.sizeof(Config) == 8 * 4 + 3 * 4 * MAX_TANKS + 8
        _lib = L
    return _lib


def make_config(mode, teams, ctrl=None, bot_types=None, map="octagon", terminate_time=None, buff_on=True,
                debuff_on=True, auto_reset=False, aim_raymarch=False, weakness=1.0):
    """teams: list of team ids per tank; ctrl: list of "agent"/"bot"; bot_types: list of names (or None)."""
    n = len(teams)
    cfg = Config()
    cfg.mode = MODE[mode] if isinstance(mode, str) else mode
    cfg.n_tanks = n
    cfg.map = MAP[map] if isinstance(map, str) else map
    cfg.terminate_time = int(terminate_time or 0)
    cfg.buff_on, cfg.debuff_on = int(buff_on), int(debuff_on)
    cfg.auto_reset, cfg.aim_raymarch = int(auto_reset), int(aim_raymarch)
    cfg.weakness = float(weakness)
    ctrl = ctrl or ["agent"] * n
    bot_types = bot_types or [None] * n
    for i in range(n):
        cfg.team[i] = int(teams[i])
        cfg.ctrl[i] = 1 if ctrl[i] == "bot" else 0
        cfg.bot_type[i] = BOT[bot_types[i]] if bot_types[i] else 0
    return cfg


def config_from_team_dict(game_configs, **kw):
    """Build a team-mode config from a reference-style dict (configs/config_teams.py)."""
    names = {}
    teams, ctrl, bots = [], [], []
    for name, tc in game_configs.items():
        if tc["mode"] == "human":
            raise ValueError("mode 'human' (keyboard) is out of scope")
        teams.append(names.setdefault(tc["team"], len(names)))
        ctrl.append("bot" if tc["mode"] == "bot" else "agent")
        bots.append(tc.get("bot_type"))
    return make_config("team", teams, ctrl, bots, **kw)


class OracleEnv:
    """N independent envs stepped on the CPU by the C oracle."""

    def __init__(self, cfg, n_envs, seed, env_id_base=0):
        self.L = lib()
        self.cfg = cfg
        self.n_envs = int(n_envs)
        self.h = self.L.ato_create(C.byref(cfg), self.n_envs, int(env_id_base), int(seed))
        self.D = self.L.ato_obs_dim(self.h)
        self.R = self.L.ato_obs_rows(self.h)
        self.A = self.L.ato_action_rows(self.h)
        self.n_tanks = cfg.n_tanks
        self.obs = np.zeros((self.n_envs, self.R * self.D), np.float32)
        self.rewards = np.zeros((self.n_envs, self.R), np.float32)
        self.done = np.zeros(self.n_envs, np.uint8)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ato_destroy(self.h)
            self.h = None

    def reset(self, mask=None):
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.ato_reset(self.h, m.ctypes.data if m is not None else None, self.obs.ctypes.data)
        return self.obs

    def step(self, actions=None, threads=0):
        a = None
        if self.A:
            a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n_envs, self.A, 3)
        ap = a.ctypes.data if a is not None else None
        if threads and threads > 1:
            self.L.ato_step_mt(self.h, ap, self.obs.ctypes.data, self.rewards.ctypes.data, self.done.ctypes.data,
                               int(threads))
        else:
            self.L.ato_step(self.h, ap, self.obs.ctypes.data, self.rewards.ctypes.data, self.done.ctypes.data)
        return self.obs, self.rewards, self.done

    def get_state(self, env=None):
        if env is None:
            out = np.zeros(self.n_envs, ENV_DT)
            for i in range(self.n_envs):
                self.L.ato_get_state(self.h, i, out[i:i + 1].ctypes.data)
            return out
        out = np.zeros(1, ENV_DT)
        self.L.ato_get_state(self.h, int(env), out.ctypes.data)
        return out[0]

    def set_state(self, env, st):
        st = np.ascontiguousarray(np.asarray(st, dtype=ENV_DT).reshape(1))
        self.L.ato_set_state(self.h, int(env), st.ctypes.data)

    def info(self):
        n = self.n_tanks
        winner = np.zeros(self.n_envs, np.int32)
        hits = np.zeros((self.n_envs, n), np.int32)
        be_hits = np.zeros((self.n_envs, n), np.int32)
        change = np.zeros((self.n_envs, n), np.int32)
        self.L.ato_get_info(self.h, winner.ctypes.data, hits.ctypes.data, be_hits.ctypes.data, change.ctypes.data)
        return {"winner": winner, "hits": hits, "be_hits": be_hits, "change_time": change}

    def terminal_info(self):
        """info of the latest step as the reference's step() returned it: for envs the harness auto-reset in that step
        the finished episode's winner / hits / be_hits (``latched`` = 1), else the current values."""
        n = self.n_tanks
        winner = np.zeros(self.n_envs, np.int32)
        hits = np.zeros((self.n_envs, n), np.int32)
        be_hits = np.zeros((self.n_envs, n), np.int32)
        change = np.zeros((self.n_envs, n), np.int32)
        latched = np.zeros(self.n_envs, np.uint8)
        self.L.ato_get_terminal_info(self.h, winner.ctypes.data, hits.ctypes.data, be_hits.ctypes.data, change.ctypes.data,
                                     latched.ctypes.data)
        return {"winner": winner, "hits": hits, "be_hits": be_hits, "change_time": change, "latched": latched}


def bfs_path(maze, start, goal):
    m = np.ascontiguousarray(np.asarray(maze, dtype=np.uint8).reshape(-1))
    out = np.zeros(2 * CELLS, np.int32)
    n = lib().ato_bfs_path(m.ctypes.data, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]), out.ctypes.data)
    if n == 0:
        return None
    return [tuple(int(v) for v in out[2 * i:2 * i + 2]) for i in range(n)]


def generate_maze(seed, env_id, ctr=0):
    m = np.zeros(CELLS, np.uint8)
    c = C.c_uint32(ctr)
    lib().ato_generate_maze(int(seed), int(env_id), C.byref(c), m.ctypes.data)
    return m.reshape(11, 11), c.value
