"""Policy / value networks and observation normalisation of the reference's PPO (models/ppo_utils.py).

Module and parameter names are the reference's (``critic.{0,2,4}``, ``actor.{h}.{0,2,4}``), so a ``state_dict`` saved
here loads into the reference's ``PPOAgentPPO`` / ``PPOAgentBot`` (inference.py:10-41, inference_multi.py:13-21) and
vice versa - tests/test_learner_cpu.py pins the key / shape list recorded from the reference classes."""
import torch
import torch.nn as nn


class RunningMeanStd:
    """models/ppo_utils.py:6-29 - Welford-style running mean / variance over the leading (batch) dimension."""

    def __init__(self, shape, epsilon=1e-4, device="cpu"):
        self.mean = torch.zeros(shape, dtype=torch.float32, device=device)
        self.var = torch.ones(shape, dtype=torch.float32, device=device)
        self.count = torch.tensor(epsilon, dtype=torch.float32, device=device)

    def update(self, x):
        x = x.to(self.mean.device)
        batch_mean = x.mean(dim=0)
        batch_var = x.var(dim=0, unbiased=False)
        batch_count = x.shape[0]
        delta = batch_mean - self.mean
        total = self.count + batch_count
        self.mean = self.mean + delta * batch_count / total
        self.var = self.var + (batch_var * batch_count + delta ** 2 * self.count * batch_count / total) / total
        self.count = total

    def normalize(self, x):
        x = x.to(self.mean.device)
        return (x - self.mean) / (torch.sqrt(self.var) + 1e-8)


def tick_normalize_cuda(obs, epsilon=1e-4, out=None):
    import ctypes as C

    from .. import _capi
    x = obs.contiguous()
    n, a, d = x.shape
    y = torch.empty_like(x) if out is None else out
    assert y.is_contiguous() and y.shape == x.shape and y.dtype == torch.float32 and x.dtype == torch.float32
    _capi.check(_capi.lib().at_tick_normalize(x.device.index, C.c_void_p(x.data_ptr()), n, a, d, float(epsilon),
                                              C.c_void_p(y.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                "at_tick_normalize")
    return y


def tick_normalize(obs, epsilon=1e-4, out=None):
    """What train_multi_ppo.py:89-91 / train_ppo_bot.py do every tick, batched over envs.

    The reference builds a FRESH ``RunningMeanStd(shape=(A, D))`` per tick, updates it with the current ``[A, D]``
    observation block and normalises that block with it; the statistics therefore run over the env's own A agent rows
    (dim 0), not over time.  ``obs`` is ``[N, A, D]``; returns the normalised ``[N, A, D]`` (same arithmetic, one env per
    batch row)."""
    if obs.is_cuda:   # one fused pass (csrc: tick_normalize_kernel) instead of ~10 elementwise / reduction kernels
        return tick_normalize_cuda(obs, epsilon, out)
    a = obs.shape[1]
    bm = obs.mean(dim=1, keepdim=True)
    bv = obs.var(dim=1, unbiased=False, keepdim=True)
    total = epsilon + a
    mean = bm * a / total                                             # 0 + delta * n / total
    var = 1.0 + (bv * a + bm ** 2 * epsilon * a / total) / total
    res = (obs - mean) / (torch.sqrt(var) + 1e-8)
    if out is not None:
        out.copy_(res)
        return out
    return res


def _mlp(obs_dim, out):
    return nn.Sequential(nn.Linear(obs_dim, 128), nn.Tanh(), nn.Linear(128, 128), nn.Tanh(), nn.Linear(128, out))


class PPOAgentPPO(nn.Module):
    """models/ppo_utils.py:31-62: one critic MLP and one actor MLP per action head ([3, 3, 2] -> rot, move, shoot)."""

    def __init__(self, obs_dim, act_dim=(3, 3, 2)):
        super().__init__()
        self.critic = _mlp(obs_dim, 1)
        self.actor = nn.ModuleList([_mlp(obs_dim, int(a)) for a in act_dim])

    def get_value(self, x):
        return self.critic(x)

    @torch.no_grad()
    def fused_logits_and_value(self, x):
        """Inference-only forward of all four MLPs as three batched GEMMs: the first layers share the input, so their
        weights are concatenated into one [4*128, D] matrix (one pass over the [N, D] observations instead of four);
        layers two and three run per network on column slices of that result (no transposed copy).  Same arithmetic per
        network as the module path."""
        nets = [self.critic] + list(self.actor)
        w1 = torch.cat([n[0].weight for n in nets], dim=0)                      # [512, D]
        b1 = torch.cat([n[0].bias for n in nets], dim=0)
        h1 = torch.addmm(b1, x, w1.t()).tanh_()                                 # [N, 512]: the four first layers side by side
        outs = []
        for k, n in enumerate(nets):    # layers two and three read their 128-column slice in place (row stride 512)
            h2 = torch.addmm(n[2].bias, h1[:, 128 * k:128 * (k + 1)], n[2].weight.t()).tanh_()
            outs.append(torch.addmm(n[4].bias, h2, n[4].weight.t()))
        return outs[1:], outs[0]

    @torch.no_grad()
    def sample(self, x):
        """(action int64 [N, 3], logprob [N, 3], value [N, 1]) through the fused forward."""
        logits, value = self.fused_logits_and_value(x)
        acts, lps = [], []
        for l in logits:
            lp = torch.log_softmax(l, dim=-1)
            a = (lp - torch.log(-torch.log(torch.rand_like(lp).clamp_min(1e-20)))).argmax(dim=-1)
            acts.append(a)
            lps.append(lp.gather(-1, a.unsqueeze(-1)).squeeze(-1))
        return torch.stack(acts, dim=-1), torch.stack(lps, dim=-1), value

    @torch.no_grad()
    def sample_cuda(self, x, actions_out, logprobs_out, counter, seed=0, stream_id=0):
        """CUDA path of ``sample``: the fused forward, then ONE kernel (csrc: sample_heads_kernel through at_sample_heads)
        for log-softmax + Gumbel-max draw + chosen log-probability of all heads, written straight into
        ``actions_out`` int32 [N, heads] / ``logprobs_out`` float32 [N, heads] (may be column slices of [N, A, heads]
        tensors).  ``counter``: device int64 tensor [1] the caller advances once per tick.  Returns the value [N, 1]."""
        import ctypes as C

        from .. import _capi
        logits, value = self.fused_logits_and_value(x)
        logits = [l.contiguous() for l in logits]
        nh = len(logits)
        assert actions_out.dtype == torch.int32 and logprobs_out.dtype == torch.float32
        assert actions_out.stride(-1) == 1 and logprobs_out.stride(-1) == 1
        ptrs = (C.c_void_p * nh)(*[l.data_ptr() for l in logits])
        dims = (C.c_int32 * nh)(*[int(l.shape[-1]) for l in logits])
        _capi.check(_capi.lib().at_sample_heads(
            x.device.index, C.cast(ptrs, C.c_void_p), C.cast(dims, C.c_void_p), nh, x.shape[0], int(seed),
            C.c_void_p(counter.data_ptr()), int(stream_id), C.c_void_p(actions_out.data_ptr()), actions_out.stride(0),
            C.c_void_p(logprobs_out.data_ptr()), logprobs_out.stride(0),
            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "at_sample_heads")
        return value

    def get_action_and_value(self, x, action=None):
        logits = [head(x) for head in self.actor]
        logp_all = [torch.log_softmax(l, dim=-1) for l in logits]
        if action is None:
            # Gumbel-max sampling == Categorical(logits).sample(), without the host sync of torch.multinomial's checks
            action = torch.stack([(lp - torch.log(-torch.log(torch.rand_like(lp).clamp_min(1e-20)))).argmax(dim=-1)
                                  for lp in logp_all], dim=-1)
        elif isinstance(action, (list, tuple)):
            action = torch.stack(list(action), dim=-1)
        idx = action.long()
        logprobs = torch.stack([lp.gather(-1, idx[..., h:h + 1]).squeeze(-1) for h, lp in enumerate(logp_all)], dim=-1)
        entropy = torch.stack([-(lp.exp() * lp).sum(dim=-1) for lp in logp_all], dim=-1)
        return action, logprobs, entropy, self.critic(x)


class PPOAgentBot(PPOAgentPPO):
    """models/ppo_utils.py:64-94 (identical architecture; the 1v1-vs-bot trainers use this name)."""
