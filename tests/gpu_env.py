"""Adapter exposing the product's BatchedTankEnv (CUDA, through the C ABI) with the env protocol of
tests/parity.py, returning numpy arrays."""
import numpy as np
import torch

from alphatank_b200 import _capi
from alphatank_b200.batched_env import BatchedTankEnv


class GpuEnv:
    def __init__(self, spec, n_envs, seed, env_id_base=0, auto_reset=False, device="cuda:0"):
        cfg = _capi.make_config(spec["mode"], spec["teams"], spec["ctrl"], spec["bots"], map=spec["map"],
                                terminate_time=spec["terminate_time"], buff_on=spec["buff_on"],
                                debuff_on=spec["debuff_on"], auto_reset=auto_reset, weakness=spec["weakness"])
        self.core = BatchedTankEnv(cfg, n_envs, device, seed, env_id_base)
        self.n_envs, self.D, self.R, self.A, self.n_tanks = n_envs, self.core.D, self.core.R, self.core.A, cfg.n_tanks

    def reset(self, mask=None):
        return self.core.reset(mask).cpu().numpy()

    def step(self, actions=None, threads=0):
        a = None if actions is None else torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32))
        o, r, d = self.core.step(a)
        return o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()

    def get_state(self, env=None):
        if env is None:
            return self.core.get_state()
        return self.core.get_state(int(env), 1)[0]

    def set_state(self, env, st):
        self.core.set_state(int(env), st)

    def info(self):
        return {k: v.cpu().numpy() for k, v in self.core.get_info().items()}

    def terminal_info(self):
        return {k: v.cpu().numpy() for k, v in self.core.get_info(terminal=True).items()}
