"""Batched PPO on the CUDA env step: the reference's train_multi_ppo.py (one PPOAgentPPO + Adam per agent tank,
GAE, clipped surrogate) re-laid for tens of thousands of envs per GPU.

* The rollout never leaves the device: ``at_step`` writes observations, cumulative rewards and done flags straight
  into the rollout storage (``obs_out`` / ``rewards_out`` / ``done_out`` of ``BatchedTankEnv.step``), actions are
  sampled on the device (Gumbel-max) and handed over as an int32 tensor.  The reference's per-tick
  tensor -> cuda -> cpu -> list hop (train_multi_ppo.py:96-116) does not exist.
* Data parallel: every rank owns a contiguous env-id range (sharding.py) and its own replicas of the A policies;
  the only collective is one all-reduce of the flattened gradient per optimiser step (NCCL on GPUs, gloo in the
  CPU tests), plus one for the advantage statistics so that every rank normalises with the global mean / std.
* Reference quirks kept behind switches (defaults are the reference's): ``reward_mode="cumulative"`` feeds the env's
  cumulative per-episode reward as the per-step reward (train_multi_ppo.py:118), ``obs_norm="tick"`` re-creates the
  observation normaliser every tick (train_multi_ppo.py:89-91).  ``"delta"`` / ``"running"`` are the conventional
  alternatives.  The reference's reset-the-env-when-any-reward-is-negative rule (:123-131) is off by default.
"""
import dataclasses
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from .ppo_utils import FusedPolicies, PPOAgentPPO, RunningMeanStd, tick_normalize


@dataclasses.dataclass
class PPOConfig:                       # defaults: train_multi_ppo.py:25-38
    learning_rate: float = 3e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_coef: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.3
    max_grad_norm: float = 0.3
    num_steps: int = 32                # ticks per rollout (the reference uses 512 with ONE env)
    num_epochs: int = 4                # (the reference uses 60 full-batch epochs on 512 samples)
    num_minibatches: int = 8
    reward_mode: str = "cumulative"    # "cumulative" (reference) | "delta"
    obs_norm: str = "tick"             # "tick" (reference) | "running" | "none"
    adam_eps: float = 1e-5


class RolloutBuffer:
    """[T(+1), N, A, ...] storage on the env's device; the env kernel writes obs / rewards / done into it directly."""

    def __init__(self, num_steps, num_envs, num_agents, obs_dim, device, keep_raw=True, keep_winner=False, obs_dtype=torch.float32):
        T, N, A, D = num_steps, num_envs, num_agents, obs_dim
        # keep_winner: info["winner"] of every tick (the finished episode's where dones[t + 1] is set; at_get_terminal_info)
        self.winner = torch.zeros((T, N), dtype=torch.int32, device=device) if keep_winner else None
        self.T, self.N, self.A, self.D = T, N, A, D
        # env output, un-normalised.  keep_raw=False: only slot 0 exists (the reset observation) - with the "tick"
        # normaliser on a GPU the env step hands over normalised rows and the raw ones are never written
        self.keep_raw = keep_raw
        self.raw_obs = torch.zeros(((T + 1) if keep_raw else 1, N, A * D), dtype=torch.float32, device=device)
        # what the policy saw (+ the bootstrap obs).  torch.float16: the env writes the normalised rows as fp16
        # (at_step_norm16) and the policies contract them on tcgen05.mma.kind::f16 - half the rollout's observation bytes
        self.obs = torch.zeros((T + 1, N, A, D), dtype=obs_dtype, device=device)
        self.actions = torch.zeros((T, N, A, 3), dtype=torch.int32, device=device)
        self.logprobs = torch.zeros((T, N, A, 3), dtype=torch.float32, device=device)
        self.env_rewards = torch.zeros((T, N, A), dtype=torch.float32, device=device)       # cumulative, as returned
        self.rewards = torch.zeros((T, N, A), dtype=torch.float32, device=device)           # what GAE consumes
        self.dones = torch.zeros((T + 1, N), dtype=torch.uint8, device=device)              # dones[t]: obs[t] starts an episode
        self.values = torch.zeros((T + 1, N, A), dtype=torch.float32, device=device)


def compute_gae(rewards, values, dones, gamma, gae_lambda):
    """train_multi_ppo.py:135-146, batched: rewards [T, ...], values [T+1, ...], dones [T+1, ...] (broadcastable).
    delta_t = r_t + gamma (1 - d_{t+1}) V_{t+1} - V_t ;  A_t = delta_t + gamma lambda (1 - d_{t+1}) A_{t+1}."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in reversed(range(T)):
        nonterminal = 1.0 - dones[t + 1].to(rewards.dtype)
        delta = rewards[t] + gamma * values[t + 1] * nonterminal - values[t]
        last = delta + gamma * gae_lambda * nonterminal * last
        adv[t] = last
    return adv


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat tensor (the PPO gradient all-reduce; no-op for a single process)."""
    w = _world()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(w)
    return flat


class PPOLearner:
    """A policies (one per agent tank, like the reference) + optimisers; device-agnostic (tested on CPU with gloo)."""

    def __init__(self, num_agents, obs_dim, act_dim=(3, 3, 2), cfg=None, device="cpu", seed=0):
        self.cfg = cfg or PPOConfig()
        self.device = torch.device(device)
        self.A, self.D, self.seed = num_agents, obs_dim, int(seed)
        g = torch.Generator().manual_seed(seed)       # identical initial replicas on every rank
        state = torch.random.get_rng_state()
        torch.random.set_rng_state(g.get_state())
        self.agents = [PPOAgentPPO(obs_dim, act_dim).to(self.device) for _ in range(num_agents)]
        torch.random.set_rng_state(state)
        self.optimizers = [torch.optim.Adam(a.parameters(), lr=self.cfg.learning_rate, eps=self.cfg.adam_eps)
                           for a in self.agents]
        self.running = [RunningMeanStd((obs_dim,), device=self.device) for _ in range(num_agents)]
        self._fused = None            # FusedPolicies per pair of agents (cuda: at_policy_forward), False: not available
        self._static = None
        self._static_keys = set()

    # ---- the tcgen05 inference path (csrc/at_policy_tc.cuh)
    def fused_policies(self, half=False):
        """The agents packed for at_policy_forward / at_policy_forward16 (two per launch), or None where that path does
        not apply (CPU, another architecture than the reference's, AT_NO_POLICY_TC=1: the per-layer kernels of
        sample_cuda take over).  ``half``: for fp16 observations."""
        if self._fused is None or (self._fused and self._fused[0].half != bool(half)):
            self._fused = False
            if self.device.type == "cuda" and not os.environ.get("AT_NO_POLICY_TC") and \
                    torch.cuda.get_device_capability(self.device)[0] == 10:
                try:
                    self._fused = [FusedPolicies(self.agents[i:i + 2], self.D, self.device, half=half) for i in range(0, self.A, 2)]
                except ValueError:
                    self._fused = False
                if self._fused and self._static is not None:
                    for k, f in enumerate(self._fused):
                        f.set_static(self._static[0], self._static[1], self._static[2][2 * k:2 * k + 2])
        return self._fused or None

    def set_static_columns(self, lo, hi, y, key=None):
        """Observation columns [lo, hi) hold y[i, lo:hi] in every row agent i sees FROM ONE ENV (BatchedTankEnv.
        static_obs_range + one normalised row): act(fold=key) on that env's rows folds them into the first layers' bias.
        Only meaningful for stateless normalisers ("tick" / "none").  ``key`` names the env the values belong to
        (collect_rollout uses id(env.core)): rows of any other env are contracted in full."""
        lo, hi = int(lo), int(hi)
        if key is not None and self._static is not None and self._static_keys:
            # a second env on the same learner (cycle trainer: one handle per bot type): its rows may use the fold only if
            # they carry the same static values - the first env's fold may be baked into a captured graph
            same = (lo, hi) == self._static[:2] and torch.equal(y[:, lo:hi].to(self._static[2].device, self._static[2].dtype),
                                                                self._static[2][:, lo:hi])
            if same:
                self._static_keys.add(key)
            return
        self._static_keys = set() if key is None else {key}
        self._static = (lo, hi, y.detach().clone())
        if self._fused:
            for k, f in enumerate(self._fused):
                f.set_static(int(lo), int(hi), self._static[2][2 * k:2 * k + 2])

    def refresh_inference(self):
        """Re-pack the inference weights after the parameters changed (update() and load_state_dict callers)."""
        if self._fused:
            for f in self._fused:
                f.refresh()

    # ---- acting
    def normalize(self, raw_obs, update=True, out=None):
        """raw_obs [N, A*D] (env layout) -> [N, A, D] as the policies see it (written into ``out`` when given)."""
        x = raw_obs.view(raw_obs.shape[0], self.A, self.D)
        mode = self.cfg.obs_norm
        if mode == "tick":
            return tick_normalize(x, out=out)
        if mode == "running":
            out = torch.empty_like(x) if out is None else out
            for i in range(self.A):
                if update:
                    self.running[i].update(x[:, i])
                out[:, i] = self.running[i].normalize(x[:, i])
            return out
        if out is not None:
            out.copy_(x)
            return out
        return x

    @torch.no_grad()
    def act(self, obs_nad, out=None, fold=None):
        """obs [N, A, D] -> actions int32 [N, A, 3], logprobs [N, A, 3], values [N, A] (written into ``out`` = the
        three tensors when given, e.g. a rollout buffer's slots).  ``fold``: the key given to set_static_columns when the
        rows come from that env (its static columns are then folded into the bias); None: contract every column."""
        if obs_nad.is_cuda:   # one sampling kernel per policy, written in place (at_sample_heads)
            n, heads = obs_nad.shape[0], len(self.agents[0].actor)
            if out is not None:
                acts, lps, vals = out
            else:
                acts = torch.empty((n, self.A, heads), dtype=torch.int32, device=obs_nad.device)
                lps = torch.empty((n, self.A, heads), dtype=torch.float32, device=obs_nad.device)
                vals = torch.empty((n, self.A), dtype=torch.float32, device=obs_nad.device)
            if getattr(self, "_sample_ctr", None) is None:     # (tick counter, the policy kernel's finished-CTA count)
                self._sample_ctr = torch.zeros(2, dtype=torch.int64, device=obs_nad.device)
            half = obs_nad.dtype == torch.float16
            fused = self.fused_policies(half)
            if fused is None and half:
                raise RuntimeError("fp16 observations need the tcgen05 inference path (at_policy_forward16)")
            do_fold = fold is not None and fold in self._static_keys and self.cfg.obs_norm in ("tick", "none")
            if fused is not None:   # the whole forward + sampling of two policies per launch (tcgen05, at_policy_forward)
                for k, f in enumerate(fused):
                    f.forward(obs_nad[:, 2 * k:2 * k + 2], acts[:, 2 * k:2 * k + 2], lps[:, 2 * k:2 * k + 2], vals[:, 2 * k:2 * k + 2],
                              self._sample_ctr, self.seed, 2 * k + self.A * _rank(), fold=do_fold,
                              advance=k == len(fused) - 1)     # the tick's last launch moves the counter on
                return acts, lps, vals
            self._sample_ctr[:1].add_(1)
            for i, agent in enumerate(self.agents):
                agent.sample_cuda(obs_nad[:, i], acts[:, i], lps[:, i], self._sample_ctr, self.seed, i + self.A * _rank(),
                                  values_out=vals[:, i])
            return acts, lps, vals
        assert out is None or not obs_nad.is_cuda
        acts, lps, vals = [], [], []
        for i, agent in enumerate(self.agents):
            a, lp, v = agent.sample(obs_nad[:, i])
            acts.append(a)
            lps.append(lp)
            vals.append(v.squeeze(-1))
        return (torch.stack(acts, dim=1).to(torch.int32), torch.stack(lps, dim=1), torch.stack(vals, dim=1))

    @torch.no_grad()
    def values_into(self, obs_nad, out, fold=None):
        """values(obs) written into ``out`` [N, A] - on the tcgen05 path one launch of the policy kernel (its draws go to
        scratch tensors and the tick counter is not advanced) instead of a cast of the rows and six GEMMs per policy."""
        fused = self.fused_policies(obs_nad.dtype == torch.float16) if obs_nad.is_cuda else None
        if fused is None:
            out.copy_(self.values(obs_nad))
            return out
        n = obs_nad.shape[0]
        sc = getattr(self, "_value_scratch", None)
        if sc is None or sc[0].shape[0] != n:
            sc = (torch.empty((n, self.A, 3), dtype=torch.int32, device=obs_nad.device),
                  torch.empty((n, self.A, 3), dtype=torch.float32, device=obs_nad.device))
            self._value_scratch = sc
        if getattr(self, "_sample_ctr", None) is None:
            self._sample_ctr = torch.zeros(2, dtype=torch.int64, device=obs_nad.device)
        do_fold = fold is not None and fold in self._static_keys and self.cfg.obs_norm in ("tick", "none")
        for k, f in enumerate(fused):
            f.forward(obs_nad[:, 2 * k:2 * k + 2], sc[0][:, 2 * k:2 * k + 2], sc[1][:, 2 * k:2 * k + 2], out[:, 2 * k:2 * k + 2],
                      self._sample_ctr, self.seed, 2 * k + self.A * _rank(), fold=do_fold)
        return out

    @torch.no_grad()
    def values(self, obs_nad):
        return torch.stack([agent.get_value(obs_nad[:, i].float()).squeeze(-1) for i, agent in enumerate(self.agents)], dim=1)

    # ---- learning
    def update(self, buf, generator=None):
        """One PPO update from a filled RolloutBuffer (train_multi_ppo.py:133-182).  Returns per-agent stats."""
        c = self.cfg
        dones = buf.dones.to(torch.float32).unsqueeze(-1)                     # [T+1, N, 1]
        adv = compute_gae(buf.rewards, buf.values, dones, c.gamma, c.gae_lambda)
        # advantage normalisation over everything (:148), with global statistics when data parallel
        s = torch.stack([adv.sum(), (adv * adv).sum(), torch.tensor(float(adv.numel()), device=adv.device)]).double()
        if _world() > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        mean = s[0] / s[2]
        var = (s[1] / s[2] - mean * mean).clamp_min(0.0) * (s[2] / (s[2] - 1.0).clamp_min(1.0))   # unbiased, like .std()
        adv = ((adv - mean) / (var.sqrt() + 1e-8)).to(torch.float32)
        returns = adv + buf.values[:-1]
        T, N = buf.T, buf.N
        B = T * N
        mb = max(B // max(c.num_minibatches, 1), 1)
        stats = []
        for i, (agent, opt) in enumerate(zip(self.agents, self.optimizers)):
            b_obs = buf.obs[:T, :, i].reshape(B, self.D)
            b_act = buf.actions[:, :, i].reshape(B, 3)
            b_lp = buf.logprobs[:, :, i].reshape(B, 3)
            b_adv = adv[:, :, i].reshape(B, 1)
            b_ret = returns[:, :, i].reshape(B)
            params = [p for p in agent.parameters()]
            last = {}
            for _ in range(c.num_epochs):
                perm = torch.randperm(B, device=b_obs.device, generator=generator) if c.num_minibatches > 1 else None
                for k in range(0, B, mb):
                    idx = perm[k:k + mb] if perm is not None else slice(None)
                    _, new_lp, entropy, new_v = agent.get_action_and_value(b_obs[idx].float(), b_act[idx])
                    ratio = torch.exp((new_lp - b_lp[idx]).clamp(-10, 10))
                    pg = torch.max(-b_adv[idx] * ratio, -b_adv[idx] * torch.clamp(ratio, 1 - c.clip_coef, 1 + c.clip_coef)).mean()
                    v_loss = 0.5 * ((new_v.squeeze(-1) - b_ret[idx]) ** 2).mean()
                    ent = entropy.mean()
                    loss = pg - c.ent_coef * ent + c.vf_coef * v_loss
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    if _world() > 1:                                          # the step path's only collective
                        flat = torch.cat([p.grad.reshape(-1) for p in params])
                        allreduce_mean_(flat)
                        o = 0
                        for p in params:
                            n = p.numel()
                            p.grad.copy_(flat[o:o + n].view_as(p))
                            o += n
                    nn.utils.clip_grad_norm_(params, c.max_grad_norm)
                    opt.step()
                    last = {"policy_loss": pg.detach(), "value_loss": v_loss.detach(), "entropy": ent.detach()}
            stats.append({k: float(v) for k, v in last.items()})
        self.refresh_inference()
        return stats


@torch.no_grad()
def collect_rollout(env, learner, buf, first=False):
    """Fill ``buf`` with ``buf.T`` ticks of ``env`` (a batched MultiAgentTeamEnv / MultiAgentEnv) under the learner's
    policies.  Everything stays on the device; the env kernel writes into the buffer.  ``first``: start from reset()."""
    core = env.core
    T = buf.T
    # GPU + "tick" normaliser: the env step returns the normalised rows itself (at_step_norm: normalised in shared
    # memory by rows_kernel), so the 277 MB of raw rows per tick are neither written nor read back
    fused = learner.cfg.obs_norm == "tick" and buf.obs.is_cuda
    half = buf.obs.dtype == torch.float16
    assert fused or buf.keep_raw, "keep_raw=False needs the fused normalised step (cuda + obs_norm='tick')"
    assert fused or not half, "fp16 observations come from the fused normalised step (cuda + obs_norm='tick')"
    if first:
        core.reset(obs_out=buf.raw_obs[0])
        if half:
            buf.obs[0].copy_(learner.normalize(buf.raw_obs[0]))       # (rounds to nearest, like at_step_norm16)
        else:
            learner.normalize(buf.raw_obs[0], out=buf.obs[0])
        buf.dones[0].zero_()
    else:                                     # continue: the previous rollout's last observation / done flag
        if buf.keep_raw:
            buf.raw_obs[0].copy_(buf.raw_obs[T])
        if fused:
            buf.obs[0].copy_(buf.obs[T])
        else:
            learner.normalize(buf.raw_obs[0], out=buf.obs[0])
        buf.dones[0].copy_(buf.dones[T])
    prev_cum = getattr(buf, "_prev_cum", None)
    if prev_cum is None or first:
        prev_cum = torch.zeros((buf.N, buf.A), dtype=torch.float32, device=buf.obs.device)
    key = id(core)
    if first and buf.obs.is_cuda and learner.cfg.obs_norm in ("tick", "none") and hasattr(core, "static_obs_range"):
        lo, hi = core.static_obs_range()        # fixed maps: the wall / maze block is the same in every row for ever
        if hi > lo:
            learner.set_static_columns(lo, hi, buf.obs[0][0], key=key)
    for t in range(T):
        if buf.obs.is_cuda:
            learner.act(buf.obs[t], out=(buf.actions[t], buf.logprobs[t], buf.values[t]), fold=key)
        else:
            a, lp, v = learner.act(buf.obs[t])
            buf.actions[t].copy_(a)
            buf.logprobs[t].copy_(lp)
            buf.values[t].copy_(v)
        if fused:
            (core.step_norm16 if half else core.step_norm)(
                buf.actions[t], buf.obs[t + 1], obs_out=buf.raw_obs[t + 1] if buf.keep_raw else None,
                rewards_out=buf.env_rewards[t], done_out=buf.dones[t + 1])
        else:
            core.step(buf.actions[t], obs_out=buf.raw_obs[t + 1], rewards_out=buf.env_rewards[t], done_out=buf.dones[t + 1])
            learner.normalize(buf.raw_obs[t + 1], update=t + 1 < T, out=buf.obs[t + 1])   # straight into the rollout storage
        if buf.winner is not None:
            core.terminal_winner(buf.winner[t])
        if learner.cfg.reward_mode == "delta":
            buf.rewards[t].copy_(buf.env_rewards[t] - prev_cum)
            prev_cum = buf.env_rewards[t] * (1.0 - buf.dones[t + 1].to(torch.float32)).unsqueeze(-1)  # reset -> 0
        else:
            buf.rewards[t].copy_(buf.env_rewards[t])
    buf._prev_cum = prev_cum
    buf._prev_live = prev_cum
    learner.values_into(buf.obs[T], buf.values[T], fold=key)
    return buf


class RolloutRunner:
    """collect_rollout() captured once as a CUDA graph and replayed: a tick is ~60 small kernels (normalise, three
    batched GEMMs per policy, sampling, the env step), and issued one by one the rollout is bound by launch overhead
    on the host, not by the GPU.  The env step takes part in the capture like any other kernel (at_step launches on
    torch's current stream).  Only for stateless observation normalisers ("tick" / "none")."""

    def __init__(self, env, learner, buf, use_graph=True):
        self.env, self.learner, self.buf = env, learner, buf
        self.use_graph = use_graph and learner.cfg.obs_norm in ("tick", "none") and buf.obs.is_cuda
        self.graph = None
        self._prev = torch.zeros((buf.N, buf.A), dtype=torch.float32, device=buf.obs.device)

    def run(self, first=False):
        buf = self.buf
        if first or not self.use_graph:
            return collect_rollout(self.env, self.learner, buf, first=first)
        if self.graph is None:
            buf._prev_cum = self._prev                      # a static tensor the captured code reads and rewrites
            self._prev.copy_(getattr(buf, "_prev_live", self._prev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                collect_rollout(self.env, self.learner, buf, first=False)
                self._prev.copy_(buf._prev_cum)
            self.graph = g                                  # (capture does not execute: replay below)
        self.graph.replay()
        return buf


# ---------------------------------------------------------------------------- checkpoints (train_multi_ppo.py:202-211)
def save_team_checkpoint(path, game_configs, tank_names, agents):
    """``{'team_config': game_configs, '<TankName>': state_dict, ...}`` - the file inference_multi.py:75-98 loads."""
    d = {"team_config": game_configs}
    for name, agent in zip(tank_names, agents):
        d[name] = {k: v.detach().cpu() for k, v in agent.state_dict().items()}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(d, path)
    return path


def load_team_checkpoint(path, obs_dim, act_dim=(3, 3, 2), device="cpu"):
    ck = torch.load(path, map_location=device, weights_only=False)
    agents = {}
    for name, sd in ck.items():
        if name == "team_config":
            continue
        a = PPOAgentPPO(obs_dim, act_dim).to(device)
        a.load_state_dict(sd)
        agents[name] = a
    return ck["team_config"], agents


# ---------------------------------------------------------------------------- 1v1 vs a scripted bot (train_ppo_bot.py)
class _OneRowCore:
    """What collect_rollout needs from ``env.core``, for ONE policy-driven tank of a 1v1 ``bot_agent`` env: the wrapper
    returns both tanks' rows and takes both tanks' action rows (gym_env.py:119, 187; the bot's row is ignored,
    train_ppo_bot.py:111-113) - this view exposes only row ``index`` of either."""

    def __init__(self, core, index):
        self.c, self.i = core, index
        self.num_envs, self.D = core.num_envs, core.D
        self._acts = torch.zeros((core.num_envs, core.A, 3), dtype=torch.int32, device=core.device)

    def reset(self, mask=None, obs_out=None):
        rows = self.c.reset(mask).view(self.num_envs, self.c.R, self.D)[:, self.i]
        if obs_out is None:
            return rows.clone()
        obs_out.view(self.num_envs, self.D).copy_(rows)
        return obs_out

    def step(self, actions, obs_out=None, rewards_out=None, done_out=None):
        self._acts[:, self.i] = actions.view(self.num_envs, 3)
        obs, rew, dn = self.c.step(self._acts, done_out=done_out)
        if obs_out is not None:
            obs_out.view(self.num_envs, self.D).copy_(obs.view(self.num_envs, self.c.R, self.D)[:, self.i])
        if rewards_out is not None:
            rewards_out.view(self.num_envs).copy_(rew[:, self.i])
        return obs_out, rewards_out, dn

    def terminal_winner(self, out):
        return self.c.terminal_winner(out)

    def step_norm(self, actions, obs_norm_out, obs_out=None, rewards_out=None, done_out=None, epsilon=1e-4):
        raw = self.c.obs.view(self.num_envs, self.c.R, self.D)
        self.step(actions, obs_out=obs_out, rewards_out=rewards_out, done_out=done_out)
        tick_normalize(raw[:, self.i:self.i + 1].contiguous(), epsilon, out=obs_norm_out.view(self.num_envs, 1, self.D))
        return obs_norm_out, rewards_out, done_out


class SingleAgentView:
    """``MultiAgentEnv(mode='bot_agent', ...)`` as a one-agent env for collect_rollout / RolloutRunner (the training agent
    is tank ``training_agent_index`` = 1, train_ppo_bot.py:35)."""

    def __init__(self, env, training_agent_index=1):
        self.env, self.core = env, _OneRowCore(env.core, training_agent_index)
