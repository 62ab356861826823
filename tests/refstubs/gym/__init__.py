"""TEST-ONLY stand-in for gym 0.26.2 (requirements.txt:2 of the reference).

Only the class skeleton the reference's wrappers subclass / instantiate
(env/gym_env.py:10,22-23; env/gym_env_multi.py:12,24-25;
models/sac_utils.py:12).  gym contributes no arithmetic to the env step.
"""
import numpy as _np
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def reset(self, **kw):
        return self.env.reset(**kw)
