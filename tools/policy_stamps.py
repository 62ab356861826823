#!/usr/bin/env python
"""Timeline of block 0 of the policy kernel (fp16 rows) (at_policy_debug_stamps): what the MMA-issuing thread and one epilogue
thread wait for, in SM clocks relative to the first stamp.  Development tooling."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphatank_b200 import _capi  # noqa: E402
from alphatank_b200.learner.ppo_utils import FusedPolicies, PPOAgentPPO  # noqa: E402

dev = torch.device("cuda:0")
N, D, A = 65536, 528, 2
torch.manual_seed(0)
agents = [PPOAgentPPO(D).to(dev) for _ in range(A)]
obs = torch.randn((N, A, D), device=dev).half()
acts = torch.zeros((N, A, 3), dtype=torch.int32, device=dev)
lps, vals = torch.zeros((N, A, 3), device=dev), torch.zeros((N, A), device=dev)
ctr = torch.tensor([1, 0], dtype=torch.int64, device=dev)
fp = FusedPolicies(agents, D, dev, half=True)
fp.set_static(237, 518, obs[0])
for _ in range(3):
    fp.forward(obs, acts, lps, vals, ctr)
st = torch.zeros((2, 512), dtype=torch.int64, device=dev)
_capi.lib().at_policy_debug_stamps(C.c_void_p(st.data_ptr()))
fp.forward(obs, acts, lps, vals, ctr)
torch.cuda.synchronize()
_capi.lib().at_policy_debug_stamps(None)
st = st.cpu().view(2, 256, 2)
names = {100: "L1 h0 begin", 101: "L1 h1 begin", 110: "L1 h0 issued", 111: "L1 h1 issued", 200: "L2 n0 begin", 201: "L2 n1 begin",
         202: "L2 n2 begin", 203: "L2 n3 begin", 210: "L2 n0 issued", 211: "L2 n1 issued", 212: "L2 n2 issued", 213: "L2 n3 issued",
         300: "epi h0: wait acc1", 301: "epi h1: wait acc1", 310: "epi h0: acc1 ready", 311: "epi h1: acc1 ready",
         320: "epi h0: convs done", 321: "epi h1: convs done", 330: "epi h0: readout a done", 331: "epi h1: readout a done",
         340: "epi h0: readout b done", 341: "epi h1: readout b done", 400: "  l2_done n0 seen", 401: "  l2_done n1 seen",
         402: "  l2_done n2 seen", 403: "  l2_done n3 seen"}
ev = []
for role in range(2):
    for i in range(256):
        e, c = int(st[role, i, 0]), int(st[role, i, 1])
        if e:
            ev.append((c, role, e))
ev.sort()
t0 = ev[0][0]
for c, role, e in ev[:140]:
    print("%8d  %s  %s" % (c - t0, "MMA" if role == 0 else "   EPI", names.get(e, str(e))))
