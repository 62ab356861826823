"""CPU: the learner-side mirrors (alphatank_b200/learner) against golden vectors recorded from the reference's own
classes (tests/golden/gen_learner_golden.py -> learner.npz), a literal restatement of the trainer's GAE loop, and
the data-parallel update over gloo (world_size 2)."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from alphatank_b200.learner import (ContinuousToDiscrete, DeviceReplayBuffer, PPOAgentBot, PPOAgentPPO, PPOConfig,  # noqa: E402
                                    PPOLearner, RolloutBuffer, RunningMeanStd, load_team_checkpoint,
                                    save_team_checkpoint, tick_normalize)
from alphatank_b200.learner.ppo import compute_gae  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "learner.npz"))


def test_state_dict_schema_is_the_reference_one():
    want = json.loads(str(G["keys_528"]))
    for cls in (PPOAgentPPO, PPOAgentBot):
        got = [[k, list(v.shape)] for k, v in cls(528, [3, 3, 2]).state_dict().items()]
        assert got == want


def test_policy_outputs_match_reference_on_its_weights():
    names = json.loads(str(G["sd_names"]))
    m = PPOAgentPPO(G["x"].shape[1], [3, 3, 2])
    m.load_state_dict({k: torch.from_numpy(G["sd/" + k]) for k in names})        # strict: same keys
    with torch.no_grad():
        a, lp, ent, val = m.get_action_and_value(torch.from_numpy(G["x"]), torch.from_numpy(G["act"]))
    assert torch.equal(a, torch.from_numpy(G["act"]))
    np.testing.assert_allclose(lp.numpy(), G["logprobs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent.numpy(), G["entropy"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(val.numpy(), G["value"], rtol=1e-5, atol=1e-6)
    # sampling: valid ranges, and the empirical distribution follows the logits
    torch.manual_seed(0)
    x = torch.from_numpy(G["x"][:1]).repeat(20000, 1)
    with torch.no_grad():
        a, lp, _, _ = m.get_action_and_value(x)
        p = torch.softmax(m.actor[0](x[:1]), -1)[0]
    assert a.shape == (20000, 3) and int(a[:, 0].max()) <= 2 and int(a[:, 2].max()) <= 1 and int(a.min()) >= 0
    freq = torch.bincount(a[:, 0], minlength=3).float() / 20000
    assert torch.allclose(freq, p, atol=0.02)


def test_tick_normalisation_is_the_trainers_fresh_running_mean_std():
    got = tick_normalize(torch.from_numpy(G["tick_in"]))
    np.testing.assert_allclose(got.numpy(), G["tick_out"], rtol=2e-5, atol=2e-5)


def test_running_mean_std_matches_reference():
    r = RunningMeanStd((G["run_in"].shape[2],))
    for b in G["run_in"]:
        r.update(torch.from_numpy(b))
    np.testing.assert_allclose(r.mean.numpy(), G["run_mean"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(r.var.numpy(), G["run_var"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(float(r.count), float(G["run_count"]), rtol=1e-7)
    np.testing.assert_allclose(r.normalize(torch.from_numpy(G["run_in"][0])).numpy(), G["run_norm"], rtol=1e-5, atol=1e-6)


def test_continuous_to_discrete_matches_wrapper_including_ties():
    got = ContinuousToDiscrete((3, 3, 2))(torch.from_numpy(G["c2d_in"]))
    assert got.dtype == torch.int32
    assert np.array_equal(got.numpy(), G["c2d_out"])


def test_gae_matches_the_trainers_loop():
    """train_multi_ppo.py:135-146 restated literally on [T, A] tensors, one env at a time."""
    torch.manual_seed(3)
    T, N, A, gamma, lam = 13, 5, 2, 0.99, 0.95
    rewards, values = torch.randn(T, N, A), torch.randn(T + 1, N, A)
    dones = (torch.rand(T + 1, N) < 0.2).float()
    got = compute_gae(rewards, values, dones.unsqueeze(-1), gamma, lam)
    for e in range(N):
        adv = torch.zeros(T, A)
        last_gae = 0
        for t in reversed(range(T)):
            next_non_terminal = 1.0 - dones[t + 1, e]
            delta = rewards[t, e] + gamma * values[t + 1, e] * next_non_terminal - values[t, e]
            last_gae = delta + gamma * lam * next_non_terminal * last_gae
            adv[t] = last_gae
        assert torch.allclose(got[:, e], adv, atol=1e-6)


def test_device_replay_buffer_ring_and_sample():
    rb = DeviceReplayBuffer(10, 3, 2, "cpu")
    s = torch.arange(8 * 3, dtype=torch.float32).view(8, 3)
    rb.push(s, torch.ones(8, 2), torch.arange(8.0), s + 1, torch.zeros(8))
    assert len(rb) == 8 and rb.position == 8
    rb.push(s[:4] + 100, torch.ones(4, 2), torch.arange(4.0) + 100, s[:4], torch.ones(4))
    assert len(rb) == 10 and rb.position == 2                      # wrapped like ReplayBuffer.push (sac_utils.py:69-73)
    assert torch.equal(rb.reward[:2], torch.tensor([102.0, 103.0])) and float(rb.reward[8]) == 100.0
    st, ac, rw, ns, dn = rb.sample(64, generator=torch.Generator().manual_seed(0))
    assert st.shape == (64, 3) and ac.shape == (64, 2) and rw.shape == (64,) and dn.shape == (64,)


def test_checkpoint_file_is_the_reference_layout(tmp_path):
    cfgs = {"Tank1": {"team": "TeamA", "color": (0, 255, 0), "mode": "agent"},
            "Tank3": {"team": "TeamB", "color": (255, 0, 0), "mode": "bot", "bot_type": "smart"}}
    L = PPOLearner(1, 20, seed=5)
    p = save_team_checkpoint(str(tmp_path / "ck" / "x.pth"), cfgs, ["Tank1"], L.agents)
    raw = torch.load(p, map_location="cpu", weights_only=False)
    assert set(raw) == {"team_config", "Tank1"} and raw["team_config"] == cfgs      # train_multi_ppo.py:205-209
    assert list(raw["Tank1"].keys()) == list(PPOAgentPPO(20).state_dict().keys())
    tc, agents = load_team_checkpoint(p, 20)
    assert tc == cfgs
    for a, b in zip(agents["Tank1"].state_dict().values(), L.agents[0].state_dict().values()):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ data parallel update over gloo
def _fake_rollout(T, N, A, D, seed):
    g = torch.Generator().manual_seed(seed)
    b = RolloutBuffer(T, N, A, D, "cpu")
    b.obs[:T].copy_(torch.randn(T, N, A, D, generator=g))
    b.actions.copy_(torch.stack([torch.randint(0, 3, (T, N, A), generator=g), torch.randint(0, 3, (T, N, A), generator=g),
                                 torch.randint(0, 2, (T, N, A), generator=g)], dim=-1).to(torch.int32))
    b.logprobs.copy_(-torch.rand(T, N, A, 3, generator=g) - 0.3)
    b.rewards.copy_(torch.randn(T, N, A, generator=g))
    b.values.copy_(torch.randn(T + 1, N, A, generator=g))
    b.dones.copy_((torch.rand(T + 1, N, generator=g) < 0.1).to(torch.uint8))
    return b


def _slice(buf, lo, hi):
    s = RolloutBuffer(buf.T, hi - lo, buf.A, buf.D, "cpu")
    for f in ("obs", "actions", "logprobs", "rewards", "values", "dones"):
        getattr(s, f).copy_(getattr(buf, f)[:, lo:hi])
    return s


CFG = dict(num_epochs=2, num_minibatches=1, num_steps=6)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _fake_rollout(6, 16, 2, 12, seed=99)
    L = PPOLearner(2, 12, cfg=PPOConfig(**CFG), seed=1)
    L.update(_slice(full, rank * 8, rank * 8 + 8))
    torch.save([a.state_dict() for a in L.agents], os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_update_equals_single_process_full_batch(tmp_path):
    def _free_port():
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s.close()
        return p
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    L = PPOLearner(2, 12, cfg=PPOConfig(**CFG), seed=1)
    L.update(_fake_rollout(6, 16, 2, 12, seed=99))
    for a0, a1, ref in zip(r0, r1, [a.state_dict() for a in L.agents]):
        for k in ref:
            assert torch.equal(a0[k], a1[k]), "replicas diverged: " + k               # same all-reduced gradient
            assert torch.allclose(a0[k], ref[k], rtol=1e-4, atol=1e-5), k              # == the full-batch update


def test_sac_networks_match_reference():
    from alphatank_b200.learner import SACAgent
    want = json.loads(str(G["sac_keys"]))
    sd = SACAgent(388, 6).state_dict()
    got = {k: ([[n, list(v.shape)] for n, v in sd[k].items()] if k != "log_alpha" else list(sd[k].shape)) for k in sd}
    assert got == want                                               # checkpoint layout of models/sac_utils.py:222-231
    ag = SACAgent(10, 3, hidden_dim=16)
    for net in ("actor", "critic1"):
        names = json.loads(str(G["sac_names/" + net]))
        getattr(ag, net).load_state_dict({k: torch.from_numpy(G["sac/%s/%s" % (net, k)]) for k in names})
    st, ac = torch.from_numpy(G["sac_state"]), torch.from_numpy(G["sac_action"])
    with torch.no_grad():
        mean, std = ag.actor(st)
        q = ag.critic1(st, ac)
        torch.manual_seed(5)
        a, lp = ag.actor.sample(st)
    for got_, key in ((mean, "sac_mean"), (std, "sac_std"), (q, "sac_q"), (a, "sac_sample_a"), (lp, "sac_sample_lp")):
        np.testing.assert_allclose(got_.numpy(), G[key], rtol=1e-5, atol=1e-6)


def test_sac_update_runs_and_moves_targets():
    from alphatank_b200.learner import SACAgent
    torch.manual_seed(0)
    ag = SACAgent(10, 3, hidden_dim=16)
    rb = DeviceReplayBuffer(256, 10, 3, "cpu")
    rb.push(torch.randn(200, 10), torch.rand(200, 3), torch.randn(200), torch.randn(200, 10), (torch.rand(200) < 0.1).float())
    before = [p.clone() for p in ag.critic1_target.parameters()]
    losses = [ag.update(rb, 64) for _ in range(3)]
    assert all(np.isfinite(v) for l in losses for v in l)
    assert any(not torch.equal(a, b) for a, b in zip(before, ag.critic1_target.parameters()))
    sd = ag.state_dict()
    ag2 = SACAgent(10, 3, hidden_dim=16)
    ag2.load_state_dict(sd)
    x = torch.randn(4, 10)
    with torch.no_grad():
        assert torch.equal(ag.actor(x)[0], ag2.actor(x)[0])
