import sys, time, torch
sys.path.insert(0, "/root/repo")
from alphatank_b200 import MultiAgentTeamEnv
from alphatank_b200.configs.config_teams import team_vs_bot_configs
env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=65536, device="cuda:0", seed=1)
env.reset()
core = env.core
torch.cuda.synchronize()
for name, fn in (("get_info (info_kernel)", lambda: core.get_info()),):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(500): fn()
    b.record(); torch.cuda.synchronize()
    print(name, "%.2f us per call" % (a.elapsed_time(b) * 1000 / 500))
x = torch.zeros(65536, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(500): x.add_(1)
b.record(); torch.cuda.synchronize()
print("torch add_ 65536 floats %.2f us per call" % (a.elapsed_time(b) * 1000 / 500))
