"""BatchedTankEnv - owner of one at_env handle (include/alphatank_b200.h) plus the torch-visible
buffers the kernels write into.  The reference-facing classes in gym_env.py / gym_env_multi.py are
thin shape adapters over this class.  PyTorch is used for device memory and streams only."""
import ctypes as C

import numpy as np
import torch

from . import _capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class LazyInfo(dict):
    """``info`` dict of step(): the device tensors are produced (one small kernel) on first access.  Like the reference's
    dict (built inside step(), gym_env.py:93-98, before the caller resets) it describes the episode that just ended
    for envs whose step returned done - also when the auto-reset has already replaced that state (at_get_terminal_info)."""

    def __init__(self, env, keys):
        super().__init__()
        self._env, self._keys, self._done = env, keys, False

    def _fill(self):
        if not self._done:
            self._done = True
            full = self._env.get_info(terminal=True)
            for k in self._keys:
                dict.__setitem__(self, k, full[k])

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        return k in self._keys

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def get(self, k, default=None):
        self._fill()
        return dict.get(self, k, default)

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)


class BatchedTankEnv:
    def __init__(self, cfg, num_envs, device="cuda:0", seed=0, env_id_base=0):
        self.L = _capi.lib()                      # raises NativeLibraryMissing when not built
        if not torch.cuda.is_available():
            raise RuntimeError("alphatank_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("alphatank_b200 envs live in GPU memory; device must be a cuda device, got %r" % (device,))
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.cfg, self.num_envs, self.seed, self.env_id_base = cfg, int(num_envs), int(seed), int(env_id_base)
        h = C.c_void_p()
        _capi.check(self.L.at_create(C.byref(cfg), self.num_envs, self.env_id_base, self.seed, self.dev_index,
                                     C.byref(h)), "at_create")
        self.h = h
        self.D = self.L.at_obs_dim(h)
        self.R = self.L.at_obs_rows(h)
        self.A = self.L.at_action_rows(h)
        self.n_tanks = cfg.n_tanks
        self.num_walls = self.L.at_num_walls(h)
        N, dev = self.num_envs, self.device
        self.obs = torch.zeros((N, self.R * self.D), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((N, self.R), dtype=torch.float32, device=dev)
        self.done = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self._actions = torch.zeros((N, max(self.A, 1), 3), dtype=torch.int32, device=dev)
        self._h_rewards = self._h_done = None
        self._tank_cols = None                    # scratch of tank_columns()

    def close(self):
        if getattr(self, "h", None):
            self.L.at_destroy(self.h)
            self.h = None

    __del__ = close

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ core calls
    def reset(self, mask=None, obs_out=None):
        obs = self.obs if obs_out is None else obs_out
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _capi.check(self.L.at_reset(self.h, _ptr(m), _ptr(obs), self._stream()), "at_reset")
        return obs

    def set_reward_mode(self, mode):
        """'cumulative': the reference's per-episode running sums (gym_env.py:186-187, gym_env_multi.py:307-308; quirk
        F1).  'delta': per-step increments - r_t - r_(t-1) inside an episode, r_t on its first step - computed inside
        the step kernel (at_set_reward_mode); a reset episode starts from zero (gaming_env.py:48-66 zeroes
        tank.reward)."""
        _capi.check(self.L.at_set_reward_mode(self.h, {"cumulative": 0, "delta": 1}[mode]), "at_set_reward_mode")

    def observe(self, obs_out=None):
        """The observation rows of the current state (env._get_observation(), gym_env.py:105-183) - nothing stepped."""
        obs = self.obs if obs_out is None else obs_out
        _capi.check(self.L.at_observe(self.h, _ptr(obs), self._stream()), "at_observe")
        return obs

    def tank_columns(self, tank):
        """float64 [7, N] on the device: alive, x, y, angle, reward, num_hit, num_be_hit of tank ``tank`` - one small
        kernel on the current stream, nothing goes through the host."""
        if self._tank_cols is None:
            self._tank_cols = torch.empty((7, self.num_envs), dtype=torch.float64, device=self.device)
        _capi.check(self.L.at_get_tank_columns(self.h, int(tank), _ptr(self._tank_cols), self._stream()),
                    "at_get_tank_columns")
        return self._tank_cols

    def _device_actions(self, actions):
        if self.A == 0:
            return None
        if isinstance(actions, torch.Tensor):
            a = actions
        else:
            a = torch.as_tensor(np.asarray(actions))
        a = a.reshape(self.num_envs, self.A, 3)
        if a.device != self.device or a.dtype != torch.int32 or not a.is_contiguous():
            self._actions.copy_(a, non_blocking=True)
            a = self._actions
        return a

    def step(self, actions=None, obs_out=None, rewards_out=None, done_out=None):
        """actions: int tensor/array [N, A, 3] (any int dtype, any device).  Outputs default to the env's own
        persistent buffers (self.obs / self.rewards / self.done) and may be redirected, e.g. into rollout storage."""
        a = self._device_actions(actions)
        obs = self.obs if obs_out is None else obs_out
        rew = self.rewards if rewards_out is None else rewards_out
        dn = self.done if done_out is None else done_out
        _capi.check(self.L.at_step(self.h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(dn), self._stream()), "at_step")
        return obs, rew, dn

    def step_norm(self, actions, obs_norm_out, obs_out=None, rewards_out=None, done_out=None, epsilon=1e-4):
        """``step`` that returns the rows tick-normalised (the reference trainers' fresh-RunningMeanStd-per-tick quirk,
        train_multi_ppo.py:89-91) into ``obs_norm_out`` - normalised while still in shared memory - and the raw rows
        only when ``obs_out`` is given.  Bit-identical to step() + learner.tick_normalize()."""
        a = self._device_actions(actions)
        rew = self.rewards if rewards_out is None else rewards_out
        dn = self.done if done_out is None else done_out
        assert obs_norm_out.is_contiguous() and obs_norm_out.numel() == self.num_envs * self.R * self.D
        _capi.check(self.L.at_step_norm(self.h, _ptr(a), _ptr(obs_out), _ptr(obs_norm_out), float(epsilon), _ptr(rew),
                                        _ptr(dn), self._stream()), "at_step_norm")
        return obs_norm_out, rew, dn

    def step_norm16(self, actions, obs_norm_out, obs_out=None, rewards_out=None, done_out=None, epsilon=1e-4):
        """``step_norm`` with the normalised rows as fp16 (``obs_norm_out``: torch.float16, N * R * D elements): the same
        fp32 values rounded to nearest - half the observation bytes, and what the tcgen05 policy kernel contracts."""
        a = self._device_actions(actions)
        rew = self.rewards if rewards_out is None else rewards_out
        dn = self.done if done_out is None else done_out
        assert obs_norm_out.dtype == torch.float16 and obs_norm_out.is_contiguous() and obs_norm_out.numel() == self.num_envs * self.R * self.D
        _capi.check(self.L.at_step_norm16(self.h, _ptr(a), _ptr(obs_out), _ptr(obs_norm_out), float(epsilon), _ptr(rew),
                                          _ptr(dn), self._stream()), "at_step_norm16")
        return obs_norm_out, rew, dn

    def step_host(self, h_actions):
        """End-to-end variant: actions come from (pinned) HOST memory, rewards and done are returned in pinned
        host memory, observations stay in HBM.  Returns as soon as rewards / done have landed in host memory (the simulation
        kernel posts a completion word the call spins on); the observations are complete in stream order, i.e. for
        anything enqueued on the current stream afterwards, without a stream synchronisation."""
        if self._h_rewards is None:
            self._h_rewards = torch.zeros((self.num_envs, self.R), dtype=torch.float32).pin_memory()
            self._h_done = torch.zeros((self.num_envs,), dtype=torch.uint8).pin_memory()
        ap, fn = None, self.L.at_step_host
        if self.A:
            if isinstance(h_actions, torch.Tensor) and h_actions.dtype == torch.uint8 and h_actions.is_contiguous() \
                    and h_actions.device.type == "cpu":
                fn = self.L.at_step_host_u8          # one byte per action: a quarter of the PCIe traffic
            elif not (isinstance(h_actions, torch.Tensor) and h_actions.dtype == torch.int32 and h_actions.is_contiguous()
                      and h_actions.device.type == "cpu"):
                h_actions = torch.as_tensor(np.asarray(h_actions), dtype=torch.int32).contiguous()
            ap = C.c_void_p(h_actions.data_ptr())
        _capi.check(fn(self.h, ap, _ptr(self.obs), C.c_void_p(self._h_rewards.data_ptr()),
                       C.c_void_p(self._h_done.data_ptr()), self._stream()), "at_step_host")
        return self.obs, self._h_rewards, self._h_done

    def set_rows_mode(self, persistent_buffers):
        """persistent_buffers=True: step calls stop rewriting the columns that never change on a fixed map (the caller's
        output buffers were written in full once before and are reused - at_set_rows_mode)."""
        _capi.check(self.L.at_set_rows_mode(self.h, 1 if persistent_buffers else 0), "at_set_rows_mode")

    def static_obs_range(self):
        """(lo, hi): observation columns [lo, hi) of every row never change while this env lives (the wall / maze
        block of a fixed map); lo == hi for per-env mazes."""
        lo, hi = C.c_int32(0), C.c_int32(0)
        _capi.check(self.L.at_obs_static_range(self.h, C.byref(lo), C.byref(hi)), "at_obs_static_range")
        return int(lo.value), int(hi.value)

    def synthetic_actions(self, tick, out=None):
        out = self._actions if out is None else out
        _capi.check(self.L.at_synthetic_actions(self.h, int(tick), _ptr(out), self._stream()), "at_synthetic_actions")
        return out

    def get_info(self, terminal=False):
        """winner / hits / be_hits / change_time of the current state; ``terminal=True``: of the latest step as the
        reference's callers see it - envs that were auto-reset in it report the episode that ended (+ ``"latched"``)."""
        N, n, dev = self.num_envs, self.n_tanks, self.device
        winner = torch.empty((N,), dtype=torch.int32, device=dev)
        hits = torch.empty((N, n), dtype=torch.int32, device=dev)
        be_hits = torch.empty((N, n), dtype=torch.int32, device=dev)
        change = torch.empty((N, n), dtype=torch.int32, device=dev)
        out = {"winner": winner, "hits": hits, "be_hits": be_hits, "change_time": change}
        if terminal:
            latched = torch.empty((N,), dtype=torch.uint8, device=dev)
            _capi.check(self.L.at_get_terminal_info(self.h, _ptr(winner), _ptr(hits), _ptr(be_hits), _ptr(change),
                                                    _ptr(latched), self._stream()), "at_get_terminal_info")
            out["latched"] = latched
        else:
            _capi.check(self.L.at_get_info(self.h, _ptr(winner), _ptr(hits), _ptr(be_hits), _ptr(change), self._stream()),
                        "at_get_info")
        return out

    def terminal_winner(self, out):
        """``info["winner"]`` of the latest step (the finished episode's where the step returned done) into ``out``
        (int32 [N]) - one small kernel, no allocation: usable inside a captured rollout."""
        assert out.dtype == torch.int32 and out.is_contiguous() and out.numel() == self.num_envs
        _capi.check(self.L.at_get_terminal_info(self.h, _ptr(out), None, None, None, None, self._stream()),
                    "at_get_terminal_info")
        return out

    def get_state(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        out = np.zeros(count, _capi.ENV_STATE_DT)
        torch.cuda.synchronize(self.device)
        _capi.check(self.L.at_get_state(self.h, int(first), int(count), out.ctypes.data), "at_get_state")
        return out

    def set_state(self, first, states):
        st = np.ascontiguousarray(np.asarray(states, dtype=_capi.ENV_STATE_DT).reshape(-1))
        torch.cuda.synchronize(self.device)
        _capi.check(self.L.at_set_state(self.h, int(first), int(st.shape[0]), st.ctypes.data), "at_set_state")

    def launch_count(self):
        return int(self.L.at_launch_count(self.h))

    def luts(self):
        trig, corn = np.zeros((361, 2)), np.zeros((361, 4))
        _capi.check(self.L.at_get_luts(self.h, trig.ctypes.data, corn.ctypes.data), "at_get_luts")
        return trig, corn


class TankView:
    """``env.game_env.tanks[i]`` stand-in: per-tank state columns as torch tensors over the batch, on the env's device
    (the reference's callers read ``tanks[i].alive`` every tick, inference.py:115-116).  Every read is one small
    column kernel (at_get_tank_columns) - no host round trip, no full-state snapshot."""

    _COLS = {"alive": 0, "x": 1, "y": 2, "angle": 3, "reward": 4, "num_hit": 5, "num_be_hit": 6}

    def __init__(self, owner, index):
        self._o, self._i = owner, index

    def _col(self, name):
        return self._o.core.tank_columns(self._i)[self._COLS[name]].clone()

    alive = property(lambda s: s._col("alive") != 0)
    x = property(lambda s: s._col("x"))
    y = property(lambda s: s._col("y"))
    angle = property(lambda s: s._col("angle").to(torch.int32))
    reward = property(lambda s: s._col("reward"))
    num_hit = property(lambda s: s._col("num_hit").to(torch.int32))
    num_be_hit = property(lambda s: s._col("num_be_hit").to(torch.int32))
