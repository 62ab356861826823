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


# ---------------------------------------------------------------------------------------------- networks / agent
import torch.nn as nn                      # noqa: E402
import torch.nn.functional as F            # noqa: E402


class QNetwork(nn.Module):
    """models/sac_utils.py:83-95 (same parameter names: fc1, fc2, fc3)."""

    def __init__(self, state_dim, action_dim, hidden_dim=256):
        super().__init__()
        self.fc1 = nn.Linear(state_dim + action_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, 1)

    def forward(self, state, action):
        x = torch.cat([state, action], dim=-1)
        return self.fc3(F.relu(self.fc2(F.relu(self.fc1(x)))))


class PolicyNetwork(nn.Module):
    """models/sac_utils.py:97-124: Gaussian policy squashed to [0, 1] by a sigmoid."""

    def __init__(self, state_dim, action_dim, hidden_dim=256, log_std_min=-20, log_std_max=2):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, action_dim)
        self.log_std = nn.Linear(hidden_dim, action_dim)
        self.log_std_min, self.log_std_max = log_std_min, log_std_max

    def forward(self, state):
        x = F.relu(self.fc2(F.relu(self.fc1(state))))
        return self.mean(x), torch.exp(torch.clamp(self.log_std(x), self.log_std_min, self.log_std_max))

    def sample(self, state):
        mean, std = self.forward(state)
        normal = torch.distributions.Normal(mean, std)
        x_t = normal.rsample()
        action = torch.sigmoid(x_t)
        log_prob = normal.log_prob(x_t) - torch.log(action * (1 - action) + 1e-6)
        return action, log_prob.sum(dim=-1, keepdim=True)


class SACAgent:
    """models/sac_utils.py:127-242, one per tank.  ``select_action`` takes a batch of observations on the device and
    returns a batch of actions on the device (the reference converts one observation from numpy and back);
    ``update`` is the reference's twin-critic update on a DeviceReplayBuffer sample.  ``state_dict`` layout as the
    reference's, so checkpoints are interchangeable."""

    def __init__(self, state_dim, action_dim, hidden_dim=256, actor_lr=3e-4, critic_lr=3e-4, alpha_lr=3e-4, gamma=0.99,
                 tau=0.005, target_entropy=None, device="cpu"):
        self.device, self.gamma, self.tau = torch.device(device), gamma, tau
        mk = lambda: QNetwork(state_dim, action_dim, hidden_dim).to(self.device)   # noqa: E731
        self.actor = PolicyNetwork(state_dim, action_dim, hidden_dim).to(self.device)
        self.critic1, self.critic2, self.critic1_target, self.critic2_target = mk(), mk(), mk(), mk()
        self.critic1_target.load_state_dict(self.critic1.state_dict())
        self.critic2_target.load_state_dict(self.critic2.state_dict())
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=actor_lr)
        self.critic1_optimizer = torch.optim.Adam(self.critic1.parameters(), lr=critic_lr)
        self.critic2_optimizer = torch.optim.Adam(self.critic2.parameters(), lr=critic_lr)
        self.target_entropy = -action_dim if target_entropy is None else target_entropy
        self.log_alpha = torch.zeros(1, requires_grad=True, device=self.device)
        self.alpha_optimizer = torch.optim.Adam([self.log_alpha], lr=alpha_lr)

    @property
    def alpha(self):
        return self.log_alpha.exp()

    @torch.no_grad()
    def select_action(self, state, evaluate=False):
        if evaluate:
            return torch.sigmoid(self.actor(state)[0])
        return self.actor.sample(state)[0]

    def update(self, replay_buffer, batch_size, generator=None):
        state, action, reward, next_state, done = replay_buffer.sample(batch_size, generator=generator)
        reward, done = reward.unsqueeze(1), done.unsqueeze(1)
        with torch.no_grad():
            next_action, next_log_prob = self.actor.sample(next_state)
            target_q = torch.min(self.critic1_target(next_state, next_action),
                                 self.critic2_target(next_state, next_action)) - self.alpha * next_log_prob
            target_q = reward + (1 - done) * self.gamma * target_q
        c1 = F.mse_loss(self.critic1(state, action), target_q)
        c2 = F.mse_loss(self.critic2(state, action), target_q)
        for opt, loss in ((self.critic1_optimizer, c1), (self.critic2_optimizer, c2)):
            opt.zero_grad()
            loss.backward()
            opt.step()
        new_action, log_prob = self.actor.sample(state)
        q_new = torch.min(self.critic1(state, new_action), self.critic2(state, new_action))
        actor_loss = (self.alpha.detach() * log_prob - q_new).mean()
        self.actor_optimizer.zero_grad()
        actor_loss.backward()
        self.actor_optimizer.step()
        alpha_loss = -(self.log_alpha * (log_prob + self.target_entropy).detach()).mean()
        self.alpha_optimizer.zero_grad()
        alpha_loss.backward()
        self.alpha_optimizer.step()
        with torch.no_grad():
            for tgt, src in ((self.critic1_target, self.critic1), (self.critic2_target, self.critic2)):
                for tp, p in zip(tgt.parameters(), src.parameters()):
                    tp.mul_(1.0 - self.tau).add_(p, alpha=self.tau)
        return tuple(float(v.detach()) for v in (actor_loss, c1, c2, alpha_loss))

    def state_dict(self):
        return {"actor": self.actor.state_dict(), "critic1": self.critic1.state_dict(), "critic2": self.critic2.state_dict(),
                "critic1_target": self.critic1_target.state_dict(), "critic2_target": self.critic2_target.state_dict(),
                "log_alpha": self.log_alpha.detach().cpu().numpy()}

    def load_state_dict(self, sd):
        for k in ("actor", "critic1", "critic2", "critic1_target", "critic2_target"):
            getattr(self, k).load_state_dict(sd[k])
        self.log_alpha = torch.tensor(sd["log_alpha"], requires_grad=True, device=self.device)
