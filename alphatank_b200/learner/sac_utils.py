"""SAC-side helpers of the reference (models/sac_utils.py) on the device.

* ``ContinuousToDiscrete`` - the action mapping of ``ContinuousToDiscreteWrapper.action`` (sac_utils.py:22-33):
  every continuous component x in [0, 1] becomes ``clip(round(x * (n - 1)), 0, n - 1)`` with numpy's round-half-to-even,
  n = [3, 3, 2]; batched as one tensor op whose int32 result feeds ``at_step`` directly.
* ``DeviceReplayBuffer`` - ``ReplayBuffer`` (sac_utils.py:63-81) as preallocated ring tensors in HBM with batched push /
  uniform sample (the reference keeps a Python list of tuples and ``random.sample``s it).
"""
import torch


class ContinuousToDiscrete:
    def __init__(self, nvec=(3, 3, 2)):
        self.nvec = tuple(int(n) for n in nvec)

    def __call__(self, action):
        """action [..., len(nvec)] float in [0, 1] -> int32 [..., len(nvec)].  torch.round is half-to-even, like
        np.round (sac_utils.py:30)."""
        n1 = torch.tensor([n - 1 for n in self.nvec], dtype=action.dtype, device=action.device)
        return torch.minimum(torch.clamp(torch.round(action * n1), min=0), n1).to(torch.int32)


class DeviceReplayBuffer:
    def __init__(self, capacity, state_dim, action_dim, device):
        self.capacity, self.device = int(capacity), torch.device(device)
        self.state = torch.zeros((capacity, state_dim), dtype=torch.float32, device=device)
        self.action = torch.zeros((capacity, action_dim), dtype=torch.float32, device=device)
        self.reward = torch.zeros((capacity,), dtype=torch.float32, device=device)
        self.next_state = torch.zeros((capacity, state_dim), dtype=torch.float32, device=device)
        self.done = torch.zeros((capacity,), dtype=torch.float32, device=device)
        self.position, self.size = 0, 0

    def push(self, state, action, reward, next_state, done):
        """Batched push of B transitions (the reference pushes one tuple per call); wraps like its ring."""
        b = state.shape[0]
        idx = (torch.arange(b, device=self.device) + self.position) % self.capacity
        self.state[idx] = state
        self.action[idx] = action.to(torch.float32)
        self.reward[idx] = reward.to(torch.float32)
        self.next_state[idx] = next_state
        self.done[idx] = done.to(torch.float32)
        self.position = (self.position + b) % self.capacity
        self.size = min(self.size + b, self.capacity)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=generator)
        return self.state[idx], self.action[idx], self.reward[idx], self.next_state[idx], self.done[idx]

    def __len__(self):
        return self.size
