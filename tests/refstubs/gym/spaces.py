"""TEST-ONLY gym.spaces subset: Box and MultiDiscrete (shape/nvec/sample)."""
import numpy as np


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
