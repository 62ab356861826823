"""Generate tests/golden/learner.npz from the UNMODIFIED reference learner classes (/root/reference/models/
ppo_utils.py and sac_utils.py, imported by path under the test-only pygame / gym stubs).  Run in the build container:

    python tests/golden/gen_learner_golden.py

Recorded: the parameter names / shapes of PPOAgentPPO(528, [3, 3, 2]) (what a reference checkpoint holds), the
outputs of a seeded small PPOAgentPPO on fixed inputs / actions, RunningMeanStd's fresh-per-tick normalisation
(train_multi_ppo.py:89-91) and three running updates, and ContinuousToDiscreteWrapper.action on a grid with ties."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "refstubs"))
sys.path.insert(0, REF)
os.chdir(REF)

from models.ppo_utils import PPOAgentPPO, RunningMeanStd  # noqa: E402
from models.sac_utils import ContinuousToDiscreteWrapper, SACAgent  # noqa: E402


def main():
    out = {}
    big = PPOAgentPPO(528, [3, 3, 2])
    out["keys_528"] = json.dumps([[k, list(v.shape)] for k, v in big.state_dict().items()])
    torch.manual_seed(7)
    D = 24
    m = PPOAgentPPO(D, [3, 3, 2])
    sd = m.state_dict()
    out["sd_names"] = json.dumps(list(sd.keys()))
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    x = torch.randn(9, D)
    act = torch.stack([torch.randint(0, 3, (9,)), torch.randint(0, 3, (9,)), torch.randint(0, 2, (9,))], dim=-1)
    with torch.no_grad():
        _, lp, ent, val = m.get_action_and_value(x, act)
    out["x"], out["act"], out["logprobs"], out["entropy"], out["value"] = x.numpy(), act.numpy(), lp.numpy(), ent.numpy(), val.numpy()
    # fresh normaliser per tick, shape (A, D), as the trainers do
    obs = torch.randn(5, 2, D) * 300 + 200          # 5 envs, A = 2
    ticks = []
    for e in range(5):
        r = RunningMeanStd(shape=(2, D))
        r.update(obs[e])
        ticks.append(r.normalize(obs[e]).numpy())
    out["tick_in"], out["tick_out"] = obs.numpy(), np.stack(ticks)
    r = RunningMeanStd(shape=(D,))
    batches = torch.randn(3, 11, D) * 50 + 10
    for b in batches:
        r.update(b)
    out["run_in"], out["run_mean"], out["run_var"], out["run_count"] = batches.numpy(), r.mean.numpy(), r.var.numpy(), r.count.numpy()
    out["run_norm"] = r.normalize(batches[0]).numpy()

    class _Env:                                      # the wrapper only reads action_space.nvec and num_tanks
        num_tanks = 2

        class action_space:
            nvec = np.array([3, 3, 2, 3, 3, 2])
            shape = (6,)
    w = ContinuousToDiscreteWrapper.__new__(ContinuousToDiscreteWrapper)
    w.env = _Env()
    grid = np.array([[a, b, c] for a in (0.0, 0.1, 0.25, 0.3, 0.5, 0.74, 0.75, 0.76, 1.0, 1.2, -0.3)
                     for b in (0.25, 0.75) for c in (0.0, 0.49, 0.5, 0.51, 1.0)], dtype=np.float32)
    out["c2d_in"] = grid
    out["c2d_out"] = np.array(w.action(grid), dtype=np.int32)
    # SAC: schema of SACAgent(388, 6).state_dict(), and a small seeded agent's deterministic + seeded-sample outputs
    big = SACAgent(388, 6).state_dict()
    out["sac_keys"] = json.dumps({k: ([[n, list(v.shape)] for n, v in big[k].items()] if k != "log_alpha" else list(big[k].shape))
                                  for k in big})
    torch.manual_seed(11)
    ag = SACAgent(10, 3, hidden_dim=16)
    sd = ag.state_dict()
    for net in ("actor", "critic1"):
        out["sac_names/" + net] = json.dumps(list(sd[net].keys()))
        for k, v in sd[net].items():
            out["sac/%s/%s" % (net, k)] = v.numpy()
    st, ac = torch.randn(7, 10), torch.rand(7, 3)
    with torch.no_grad():
        mean, std = ag.actor(st)
        q = ag.critic1(st, ac)
        torch.manual_seed(5)
        a, lp = ag.actor.sample(st)
    out["sac_state"], out["sac_action"], out["sac_mean"], out["sac_std"], out["sac_q"] = st.numpy(), ac.numpy(), mean.numpy(), std.numpy(), q.numpy()
    out["sac_sample_a"], out["sac_sample_lp"] = a.numpy(), lp.numpy()
    np.savez_compressed(os.path.join(HERE, "learner.npz"), **out)
    print("wrote learner.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
