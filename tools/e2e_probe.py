import sys, time, ctypes as C, torch
sys.path.insert(0, "/root/repo")
from alphatank_b200 import MultiAgentTeamEnv
from alphatank_b200.configs.config_teams import team_vs_bot_configs
E = 65536
env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=E, device="cuda:0", seed=0, map="octagon", terminate_time=None, auto_reset=True)
core = env.core
env.reset()
acts = torch.empty((E, 2, 3), dtype=torch.int32, device="cuda")
core.synthetic_actions(0, acts)
h = acts.cpu().pin_memory()
for t in range(20): env.step_host(h)
def wall(fn, n=200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("step_host (python wrapper)      %.4f ms" % wall(lambda: env.step_host(h)))
L = core.L; hdl = core.h
ap = C.c_void_p(h.data_ptr()); op = C.c_void_p(core.obs.data_ptr()); rp = C.c_void_p(core._h_rewards.data_ptr()); dp = C.c_void_p(core._h_done.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
print("at_step_host (ctypes only)      %.4f ms" % wall(lambda: L.at_step_host(hdl, ap, op, rp, dp, st)))
print("core.step + sync (device acts)  %.4f ms" % wall(lambda: (core.step(acts), torch.cuda.synchronize())))
print("core.step async back-to-back    %.4f ms" % wall(lambda: core.step(acts)))
x = torch.zeros(1, device="cuda")
print("tiny torch kernel + sync        %.4f ms" % wall(lambda: (x.add_(1), torch.cuda.synchronize())))
