"""Minimal Box / MultiDiscrete with the attributes the reference's callers read
(``observation_space.shape[0]``, ``action_space.nvec[:3]``, ``action_space.sample()``;
train_ppo_bot.py:49-50, train_multi_ppo.py:47-48).  gym is used instead when it is importable."""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym.spaces import Box, MultiDiscrete  # type: ignore
except Exception:  # noqa: BLE001
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def sample(self):
            return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)

    class MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape = self.nvec.shape

        def sample(self):
            return (np.random.random(self.nvec.shape) * self.nvec).astype(np.int64)
