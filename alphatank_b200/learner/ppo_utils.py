"""Policy / value networks and observation normalisation of the reference's PPO (models/ppo_utils.py).

Module and parameter names are the reference's (``critic.{0,2,4}``, ``actor.{h}.{0,2,4}``), so a ``state_dict`` saved
here loads into the reference's ``PPOAgentPPO`` / ``PPOAgentBot`` (inference.py:10-41, inference_multi.py:13-21) and
vice versa - tests/test_learner_cpu.py pins the key / shape list recorded from the reference classes."""
import os

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
    def sample_cuda(self, x, actions_out, logprobs_out, counter, seed=0, stream_id=0, values_out=None):
        """CUDA path of ``sample``, written straight into ``actions_out`` int32 [N, heads] / ``logprobs_out`` float32
        [N, heads] / ``values_out`` float32 [N] (may be column slices of [N, A, ...] tensors).  ``counter``: device int64
        tensor [1] the caller advances once per tick.  Returns the value [N, 1].

        Reference architecture (act_dim (3, 3, 2), 128 hidden units): the first layers as one GEMM, the second layers'
        pre-activations as four GEMMs on column slices, then ONE kernel (csrc: policy_tail_kernel through
        at_policy_tail) for tanh + the four output layers + log-softmax + Gumbel-max draw + chosen log-probability.
        Anything else: the fused forward, then at_sample_heads."""
        import ctypes as C

        from .. import _capi
        n = x.shape[0]
        nets = [self.critic] + list(self.actor)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        assert actions_out.dtype == torch.int32 and logprobs_out.dtype == torch.float32
        assert actions_out.stride(-1) == 1 and logprobs_out.stride(-1) == 1
        if [int(m[4].out_features) for m in nets] == [1, 3, 3, 2] and all(m[2].out_features == 128 for m in nets):
            w1 = torch.cat([m[0].weight for m in nets], dim=0)                      # [512, D]
            b1 = torch.cat([m[0].bias for m in nets], dim=0)
            w3 = torch.cat([m[4].weight for m in nets], dim=0).contiguous()         # [9, 128]
            b3 = torch.cat([m[4].bias for m in nets], dim=0).contiguous()
            value = torch.empty((n, 1), dtype=torch.float32, device=x.device) if values_out is None else values_out
            if all(m[0].out_features == 128 for m in nets) and not os.environ.get("AT_NO_POLICY_MID"):
                # first layers by cuBLAS, everything behind them in two kernels (csrc: policy_mid_kernel - tanh, the
                # second layers on the tensor cores, tanh, output layers - and sample_out9_kernel)
                z1 = torch.addmm(b1, x, w1.t())                                     # [N, 512] pre-activations
                w2 = torch.stack([m[2].weight for m in nets]).contiguous()          # [4, 128 out, 128 in]
                b2 = torch.stack([m[2].bias for m in nets]).contiguous()
                out9 = torch.empty((n, 9), dtype=torch.float32, device=x.device)
                _capi.check(_capi.lib().at_policy_mid_tail(
                    x.device.index, C.c_void_p(z1.data_ptr()), C.c_void_p(w2.data_ptr()), C.c_void_p(b2.data_ptr()),
                    C.c_void_p(w3.data_ptr()), C.c_void_p(b3.data_ptr()), C.c_void_p(out9.data_ptr()), n, int(seed),
                    C.c_void_p(counter.data_ptr()), int(stream_id), C.c_void_p(actions_out.data_ptr()), actions_out.stride(0),
                    C.c_void_p(logprobs_out.data_ptr()), logprobs_out.stride(0), C.c_void_p(value.data_ptr()), value.stride(0),
                    stream), "at_policy_mid_tail")
                return value.view(n, 1) if values_out is None else value
            h1 = torch.addmm(b1, x, w1.t()).tanh_()                                 # [N, 512]
            z2 = [torch.addmm(m[2].bias, h1[:, 128 * k:128 * (k + 1)], m[2].weight.t()) for k, m in enumerate(nets)]
            ptrs = (C.c_void_p * 4)(*[z.data_ptr() for z in z2])
            _capi.check(_capi.lib().at_policy_tail(
                x.device.index, C.cast(ptrs, C.c_void_p), C.c_void_p(w3.data_ptr()), C.c_void_p(b3.data_ptr()), n, int(seed),
                C.c_void_p(counter.data_ptr()), int(stream_id), C.c_void_p(actions_out.data_ptr()), actions_out.stride(0),
                C.c_void_p(logprobs_out.data_ptr()), logprobs_out.stride(0), C.c_void_p(value.data_ptr()), value.stride(0),
                stream), "at_policy_tail")
            return value.view(n, 1) if values_out is None else value
        logits, value = self.fused_logits_and_value(x)
        logits = [l.contiguous() for l in logits]
        nh = len(logits)
        ptrs = (C.c_void_p * nh)(*[l.data_ptr() for l in logits])
        dims = (C.c_int32 * nh)(*[int(l.shape[-1]) for l in logits])
        _capi.check(_capi.lib().at_sample_heads(
            x.device.index, C.cast(ptrs, C.c_void_p), C.cast(dims, C.c_void_p), nh, n, int(seed),
            C.c_void_p(counter.data_ptr()), int(stream_id), C.c_void_p(actions_out.data_ptr()), actions_out.stride(0),
            C.c_void_p(logprobs_out.data_ptr()), logprobs_out.stride(0), stream), "at_sample_heads")
        if values_out is not None:
            values_out.copy_(value.view(values_out.shape))
            return values_out
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


def _round_tf32_(w):
    """Round an fp32 tensor to the nearest TF32 value in place (10-bit mantissa): the tensor core truncates its fp32
    operands, rounding the weights once on the host keeps their error unbiased."""
    bits = w.view(torch.int32)
    bits.add_(0x1000).bitwise_and_(~0x1FFF)
    return w


class FusedPolicies:
    """The inference side of up to two PPOAgentPPO policies on the tcgen05 tensor cores (csrc/at_policy_tc.cuh through
    at_policy_forward): one kernel launch samples actions / log-probabilities / values for all of them.  The weights
    are packed once into static tensors (``refresh()`` after every optimiser step re-packs them in place, so a
    captured CUDA graph keeps reading the right storage).

    ``static``: (lo, hi, y) - observation columns [lo, hi) are the same for every row the policies will ever see and
    hold the values y[i, lo:hi] for agent i (the wall / maze block of a fixed map after the stateless per-tick
    normaliser).  K chunks of 32 columns that lie inside that range are dropped from the first layers' contraction and
    their contribution W1[:, chunk] . y is folded into the bias."""

    def __init__(self, agents, obs_dim, device, half=False):
        assert 1 <= len(agents) <= 2
        self.agents, self.D, self.device = list(agents), int(obs_dim), torch.device(device)
        # half: fp16 observations (at_step_norm16) and fp16 weights on tcgen05.mma.kind::f16 (at_policy_forward16) instead
        # of fp32 observations / TF32 (at_policy_forward); K chunks are 64 columns wide there
        self.half, self.kc = bool(half), 64 if half else 32
        for a in self.agents:
            nets = [a.critic] + list(a.actor)
            if [int(m[4].out_features) for m in nets] != [1, 3, 3, 2] or any(m[0].out_features != 128 or m[2].out_features != 128 for m in nets):
                raise ValueError("at_policy_forward serves the reference architecture only (128 hidden units, heads 3 / 3 / 2)")
        if self.D % (8 if self.half else 4) != 0 or self.D > 1024:     # (TMA: 16-byte row pitch)
            raise ValueError("at_policy_forward needs an observation width that is a multiple of 4 (fp16: 8), up to 1024")
        n = len(self.agents)
        f = dict(dtype=torch.float32, device=self.device)
        wt = torch.float16 if self.half else torch.float32
        self.w1 = torch.empty((n, 512, self.D), dtype=wt, device=self.device)
        self.b1 = torch.empty((n, 512), **f)
        self.b1_fold = torch.empty((n, 512), **f)        # b1 + W1[:, dropped chunks] . y (forward(fold=True))
        self.w2 = torch.empty((n, 4, 128, 128), dtype=torch.float16, device=self.device)   # (fp16 in both kernels)
        self.b2 = torch.empty((n, 4, 128), **f)
        self.w3 = torch.empty((n, 9, 128), **f)
        self.b3 = torch.empty((n, 9), **f)
        self.chunk_mask = 0          # 0 = every chunk
        self._fold_cols = None
        self._fold_y = None
        self.refresh()

    def set_static(self, lo, hi, y):
        """y: [n_agents, D] the constant columns' values as the policies see them (only [lo, hi) is read)."""
        kc = self.kc
        k_chunks = (self.D + kc - 1) // kc
        keep = [c for c in range(k_chunks) if not (lo <= kc * c and min(kc * c + kc, self.D) <= hi)]
        if len(keep) == k_chunks:
            self.chunk_mask, self._fold_cols, self._fold_y = 0, None, None
        else:
            drop = [c for c in range(k_chunks) if c not in keep]
            cols = torch.cat([torch.arange(kc * c, min(kc * c + kc, self.D)) for c in drop]).to(self.device)
            self.chunk_mask = sum(1 << c for c in keep)
            self._fold_cols = cols
            self._fold_y = y.to(self.device, torch.float32)[:, cols].contiguous()
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        for i, a in enumerate(self.agents):
            nets = [a.critic] + list(a.actor)
            w1 = torch.cat([m[0].weight for m in nets], dim=0)
            b1 = torch.cat([m[0].bias for m in nets], dim=0)
            self.w1[i].copy_(w1)
            self.b1[i].copy_(b1)
            if self._fold_cols is not None:      # fp32 fold of the dropped columns (exact weights, before the TF32 rounding)
                b1 = b1 + w1[:, self._fold_cols] @ self._fold_y[i]
            self.b1_fold[i].copy_(b1)
            self.w2[i].copy_(torch.stack([m[2].weight for m in nets]))
            self.b2[i].copy_(torch.stack([m[2].bias for m in nets]))
            self.w3[i].copy_(torch.cat([m[4].weight for m in nets], dim=0))
            self.b3[i].copy_(torch.cat([m[4].bias for m in nets], dim=0))
        if not self.half:         # (copy_ into the fp16 tensors has rounded to nearest already)
            _round_tf32_(self.w1)

    def forward(self, obs_nad, actions_out, logprobs_out, values_out, counter, seed=0, stream_base=0, out9=None, fold=None, advance=False):
        """obs [N, A, D] (A = the number of packed policies) -> actions int32 [N, A, 3], logprobs [N, A, 3], values [N, A]
        written in place; ``counter``: device int64 [2] = (tick counter, 0); ``advance``: the kernel moves the tick counter on
        when it is done (else the caller does, once per tick); ``out9``: optional
        [A, N, 9] (value + 8 logits per row).  ``fold``: contract only the K chunks kept by set_static() (the rows MUST hold
        the static values given there); default: whenever set_static() dropped chunks."""
        fold = bool(self.chunk_mask) if fold is None else (bool(fold) and bool(self.chunk_mask))
        import ctypes as C

        from .. import _capi
        n, A = obs_nad.shape[0], len(self.agents)
        assert obs_nad.shape[1] == A and obs_nad.shape[2] == self.D and obs_nad.stride(2) == 1
        assert obs_nad.dtype == (torch.float16 if self.half else torch.float32)
        assert actions_out.dtype == torch.int32 and actions_out.stride(-1) == 1 and logprobs_out.stride(-1) == 1
        pols = (_capi.AtPolicy * A)()
        for i in range(A):
            x, a, lp, v = obs_nad[:, i], actions_out[:, i], logprobs_out[:, i], values_out[:, i]
            q = pols[i]
            q.x, q.x_row_stride = x.data_ptr(), x.stride(0)
            q.w1, q.b1 = self.w1[i].data_ptr(), (self.b1_fold if fold else self.b1)[i].data_ptr()
            q.w2, q.b2 = self.w2[i].data_ptr(), self.b2[i].data_ptr()
            q.w3, q.b3 = self.w3[i].data_ptr(), self.b3[i].data_ptr()
            q.actions, q.act_stride = a.data_ptr(), a.stride(0)
            q.logprobs, q.lp_stride = lp.data_ptr(), lp.stride(0)
            q.values, q.val_stride = v.data_ptr(), v.stride(0)
            q.out9 = out9[i].data_ptr() if out9 is not None else None
            q.stream_id = int(stream_base) + i
        fn = _capi.lib().at_policy_forward16 if self.half else _capi.lib().at_policy_forward
        assert counter.dtype == torch.int64 and counter.numel() >= 2
        _capi.check(fn(obs_nad.device.index, pols, A, n, self.D, int(self.chunk_mask) if fold else 0, int(seed), C.c_void_p(counter.data_ptr()),
                       1 if advance else 0, C.c_void_p(torch.cuda.current_stream(obs_nad.device).cuda_stream)), "at_policy_forward")


class PPOAgentBot(PPOAgentPPO):
    """models/ppo_utils.py:64-94 (identical architecture; the 1v1-vs-bot trainers use this name)."""
