"""ctypes binding of the C ABI in include/alphatank_b200.h (libalphatank_b200.so, built in-tree by
__graft_entry__.build()).  This is the only place the package touches native code.  There is no CPU
fallback: importing works anywhere, but creating an env without the built library or without an sm_100
GPU raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libalphatank_b200.so")

MAX_TANKS, MAX_BULLETS, CELLS = 8, 48, 121
MODE = {"agent": 0, "bot_agent": 1, "arena": 2, "team": 3}
BOT_TYPES = {"smart": 0, "random": 1, "aggressive": 2, "defensive": 3, "dodge": 4, "stationary": 5}
MAP = {"octagon": 0, "battlefield": 1, "maze": 2}
AT_OK, AT_ERR_ARG, AT_ERR_CUDA, AT_ERR_NOGPU = 0, -1, -2, -3


class AtConfig(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_tanks", C.c_int32), ("map", C.c_int32), ("terminate_time", C.c_int32),
                ("buff_on", C.c_int32), ("debuff_on", C.c_int32), ("auto_reset", C.c_int32), ("reserved", C.c_int32),
                ("team", C.c_int32 * MAX_TANKS), ("ctrl", C.c_int32 * MAX_TANKS), ("bot_type", C.c_int32 * MAX_TANKS),
                ("weakness", C.c_double)]


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
ENV_STATE_DT = np.dtype([("n_bullets", "i4"), ("tick", "i4"), ("training_step", "i4"), ("episode_steps", "i4"),
                         ("rng_ctr", "u4"), ("n_walls", "i4"), ("zones", "i4", (8,)), ("maze", "u1", (CELLS + 7,)),
                         ("tanks", TANK_DT, (MAX_TANKS,)), ("bots", BOT_DT, (MAX_TANKS,)),
                         ("bullets", BULLET_DT, (MAX_BULLETS,))], align=True)

EXPORTS = ("at_create at_destroy at_obs_dim at_obs_rows at_action_rows at_num_walls at_num_envs at_reset at_step at_step_norm "
           "at_step_host at_get_info at_get_terminal_info at_get_state at_set_state at_synthetic_actions at_bfs_batch at_generate_mazes "
           "at_get_luts at_tick_normalize at_sample_heads at_policy_tail at_policy_mid_tail at_debug_block_times at_set_weakness at_launch_count at_last_error at_version "
           "at_observe at_set_reward_mode at_get_tank_columns at_forget_host_buffers at_policy_first at_policy_forward at_obs_static_range at_step_norm16 at_policy_forward16 at_policy_debug_stamps at_set_rows_mode at_step_host_u8").split()

_lib = None


class AtPolicy(C.Structure):
    """include/alphatank_b200.h: at_policy (one policy's operands and outputs for at_policy_forward)."""
    _fields_ = [("x", C.c_void_p), ("x_row_stride", C.c_int64),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
                ("actions", C.c_void_p), ("act_stride", C.c_int64), ("logprobs", C.c_void_p), ("lp_stride", C.c_int64),
                ("values", C.c_void_p), ("val_stride", C.c_int64), ("out9", C.c_void_p), ("stream_id", C.c_uint32)]


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load libalphatank_b200.so (once).  Raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "alphatank_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.at_create.argtypes = [C.POINTER(AtConfig), i64, i64, u64, i32, C.POINTER(vp)]
    L.at_destroy.argtypes = [vp]
    for f in ("at_obs_dim", "at_obs_rows", "at_action_rows", "at_num_walls"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = i32
    for f in ("at_num_envs", "at_launch_count"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = i64
    L.at_reset.argtypes = [vp, vp, vp, vp]
    L.at_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.at_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.at_step_host_u8.argtypes = [vp, vp, vp, vp, vp, vp]
    L.at_step_norm.argtypes = [vp, vp, vp, vp, C.c_float, vp, vp, vp]
    L.at_step_norm16.argtypes = [vp, vp, vp, vp, C.c_float, vp, vp, vp]
    L.at_get_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.at_get_terminal_info.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.at_get_state.argtypes = [vp, i64, i64, vp]
    L.at_set_state.argtypes = [vp, i64, i64, vp]
    L.at_synthetic_actions.argtypes = [vp, i64, vp, vp]
    L.at_bfs_batch.argtypes = [i32, vp, vp, i64, vp, vp]
    L.at_generate_mazes.argtypes = [i32, u64, i64, i64, vp, vp, vp]
    L.at_get_luts.argtypes = [vp, vp, vp]
    L.at_set_weakness.argtypes = [vp, C.c_double]
    L.at_observe.argtypes = [vp, vp, vp]
    L.at_set_reward_mode.argtypes = [vp, i32]
    L.at_get_tank_columns.argtypes = [vp, i32, vp, vp]
    L.at_forget_host_buffers.argtypes = [vp]
    L.at_policy_first.argtypes = [i32, vp, i64, i64, i32, vp, vp, vp, vp]
    L.at_policy_forward.argtypes = [i32, C.POINTER(AtPolicy), i32, i64, i32, C.c_uint32, u64, vp, i32, vp]
    L.at_policy_forward16.argtypes = [i32, C.POINTER(AtPolicy), i32, i64, i32, C.c_uint32, u64, vp, i32, vp]
    L.at_policy_debug_stamps.argtypes = [vp]
    L.at_set_rows_mode.argtypes = [vp, i32]
    L.at_obs_static_range.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.at_tick_normalize.argtypes = [i32, vp, i64, i32, i32, C.c_float, vp, vp]
    L.at_debug_block_times.argtypes = [vp, vp]
    L.at_policy_mid_tail.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, C.c_uint64, vp, C.c_uint32, vp, i64, vp, i64, vp, i64, vp]
    L.at_policy_tail.argtypes = [i32, vp, vp, vp, i64, C.c_uint64, vp, C.c_uint32, vp, i64, vp, i64, vp, i64, vp]
    L.at_sample_heads.argtypes = [i32, vp, vp, i32, i64, C.c_uint64, vp, C.c_uint32, vp, i64, vp, i64, vp]
    L.at_last_error.restype = C.c_char_p
    L.at_version.restype = C.c_char_p
    assert C.sizeof(AtConfig) == 8 * 4 + 3 * 4 * MAX_TANKS + 8
    _lib = L
    return L


def check(rc, what):
    if rc != AT_OK:
        msg = lib().at_last_error().decode(errors="replace")
        if rc == AT_ERR_ARG:
            raise ValueError("%s: %s" % (what, msg))
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


def make_config(mode, teams, ctrl, bot_types, map="octagon", terminate_time=None, buff_on=True, debuff_on=True,
                auto_reset=True, weakness=1.0):
    n = len(teams)
    if not 2 <= n <= MAX_TANKS:
        raise ValueError("between 2 and %d tanks are supported, got %d" % (MAX_TANKS, n))
    cfg = AtConfig()
    cfg.mode = MODE[mode]
    cfg.n_tanks = n
    if map not in MAP:
        raise ValueError("unknown map %r (valid: %s)" % (map, sorted(MAP)))
    cfg.map = MAP[map]
    cfg.terminate_time = int(terminate_time or 0)
    cfg.buff_on, cfg.debuff_on, cfg.auto_reset = int(bool(buff_on)), int(bool(debuff_on)), int(bool(auto_reset))
    cfg.weakness = float(weakness)
    for i in range(n):
        cfg.team[i] = int(teams[i])
        cfg.ctrl[i] = 1 if ctrl[i] == "bot" else 0
        bt = bot_types[i]
        if ctrl[i] == "bot":
            if bt not in BOT_TYPES:  # same error as BotFactory.create_bot (env/bots/bot_factory.py:29-31)
                raise ValueError("Unknown bot type: %s. Valid types are: %s" % (bt, list(BOT_TYPES.keys())))
            cfg.bot_type[i] = BOT_TYPES[bt]
    return cfg
