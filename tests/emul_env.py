"""TEST-ONLY ctypes wrapper of tests/cpu_emul/libat_emul.so (host emulation of the CUDA kernel's per-thread
code, see emul.cpp).  Implements the env protocol of tests/parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_emul", "emul.cpp")
LIB = os.path.join(HERE, "cpu_emul", "libat_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "alphatank_b200", "csrc")


def build():
    deps = [SRC, os.path.join(HERE, "cpu_emul", "simt.h")] + [os.path.join(CSRC, f) for f in ("at_sim.h", "at_host.h", "at_types.h", "at_rows.h", "at_warp.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-ffp-contract=off", "-mfma", "-std=c++17", "-shared", "-o", LIB,
                               SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.em_create.argtypes = [vp, i64, i64, C.c_uint64, C.POINTER(vp)]
        L.em_destroy.argtypes = [vp]
        for f in ("em_obs_dim", "em_obs_rows", "em_action_rows"):
            getattr(L, f).argtypes = [vp]
        L.em_reset.argtypes = [vp, vp, vp]
        L.em_set_lane_order.argtypes = [i32]
        L.em_set_weakness.argtypes = [vp, C.c_double]
        L.em_collectives.argtypes = [vp]
        L.em_collectives.restype = C.c_uint64
        L.em_set_spec_cap.argtypes = [i32]
        L.em_step.argtypes = [vp, vp, vp, vp, vp]
        L.em_get_state.argtypes = [vp, i64, vp]
        L.em_set_state.argtypes = [vp, i64, vp]
        L.em_get_info.argtypes = [vp, vp, vp, vp, vp]
        L.em_get_terminal_info.argtypes = [vp, vp, vp, vp, vp, vp]
        L.em_bfs.argtypes = [vp, i32, i32, i32, i32, vp, vp]
        L.em_generate_maze.argtypes = [C.c_uint64, i64, vp, vp]
        L.em_jump_axis.restype = C.c_double
        L.em_jump_axis.argtypes = [C.c_double, C.c_double, i32]
        L.em_atan2_refine.restype = C.c_double
        L.em_atan2_refine.argtypes = [C.c_double, C.c_double, i32]
        L.em_last_error.restype = C.c_char_p
        _lib = L
    return _lib


class EmulEnv:
    def __init__(self, cfg, n_envs, seed, env_id_base=0):
        self.L = lib()
        self.n_envs = int(n_envs)
        h = C.c_void_p()
        rc = self.L.em_create(C.byref(cfg), self.n_envs, int(env_id_base), int(seed), C.byref(h))
        if rc != 0:
            raise ValueError(self.L.em_last_error().decode())
        self.h = h
        self.D, self.R, self.A = self.L.em_obs_dim(h), self.L.em_obs_rows(h), self.L.em_action_rows(h)
        self.n_tanks = cfg.n_tanks
        self.obs = np.zeros((self.n_envs, self.R * self.D), np.float32)
        self.rewards = np.zeros((self.n_envs, self.R), np.float32)
        self.done = np.zeros(self.n_envs, np.uint8)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.em_destroy(self.h)
            self.h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.em_reset(self.h, m.ctypes.data if m is not None else None, self.obs.ctypes.data)
        return self.obs

    def step(self, actions=None, threads=0):
        a = None
        if self.A:
            a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n_envs, self.A, 3)
        self.L.em_step(self.h, a.ctypes.data if a is not None else None, self.obs.ctypes.data,
                       self.rewards.ctypes.data, self.done.ctypes.data)
        return self.obs, self.rewards, self.done

    def get_state(self, env=None):
        if env is None:
            out = np.zeros(self.n_envs, O.ENV_DT)
            for i in range(self.n_envs):
                self.L.em_get_state(self.h, i, out[i:i + 1].ctypes.data)
            return out
        out = np.zeros(1, O.ENV_DT)
        self.L.em_get_state(self.h, int(env), out.ctypes.data)
        return out[0]

    def set_state(self, env, st):
        st = np.ascontiguousarray(np.asarray(st, dtype=O.ENV_DT).reshape(1))
        rc = self.L.em_set_state(self.h, int(env), st.ctypes.data)
        if rc != 0:
            raise ValueError(self.L.em_last_error().decode())

    def info(self):
        n = self.n_tanks
        winner = np.zeros(self.n_envs, np.int32)
        hits = np.zeros((self.n_envs, n), np.int32)
        be_hits = np.zeros((self.n_envs, n), np.int32)
        change = np.zeros((self.n_envs, n), np.int32)
        self.L.em_get_info(self.h, winner.ctypes.data, hits.ctypes.data, be_hits.ctypes.data, change.ctypes.data)
        return {"winner": winner, "hits": hits, "be_hits": be_hits, "change_time": change}

    def terminal_info(self):
        n = self.n_tanks
        winner = np.zeros(self.n_envs, np.int32)
        hits = np.zeros((self.n_envs, n), np.int32)
        be_hits = np.zeros((self.n_envs, n), np.int32)
        change = np.zeros((self.n_envs, n), np.int32)
        latched = np.zeros(self.n_envs, np.uint8)
        self.L.em_get_terminal_info(self.h, winner.ctypes.data, hits.ctypes.data, be_hits.ctypes.data, change.ctypes.data,
                                    latched.ctypes.data)
        return {"winner": winner, "hits": hits, "be_hits": be_hits, "change_time": change, "latched": latched}


def make_emul_env(spec, n_envs, seed, env_id_base=0, auto_reset=False):
    cfg = O.make_config(spec["mode"], spec["teams"], spec["ctrl"], spec["bots"], map=spec["map"],
                        terminate_time=spec["terminate_time"], buff_on=spec["buff_on"], debuff_on=spec["debuff_on"],
                        auto_reset=auto_reset, weakness=spec["weakness"])
    return EmulEnv(cfg, n_envs, seed, env_id_base)
