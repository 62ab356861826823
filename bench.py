#!/usr/bin/env python
"""bench.py - agent-steps/s of the batched alphaTank environment step on N B200s of one node.

Workload (BASELINE.json configs[2]): 2v2 ``team_vs_bot_configs`` (Tank1,2 agents vs Tank3,4 on-device smart
bots), octagon map, 65536 envs per GPU, synthetic uniform random actions (ATRNG stream), auto-reset.
One "step" = one tick of every env = ONE fused kernel launch per GPU.  Envs shard across GPUs by env id with
no data-path collective (weak scaling: per-GPU work fixed).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 --steps 200 --warmup 10
    python bench.py --impl reference --steps 20 --warmup 3      # CPU arm: oracle port on all host threads

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "agent_steps_per_sec", "agent-steps/s"
ALG_BYTES_PER_ENV_STEP = 7039      # 2*S_state(1400) + 3*A + 4*R*D + 4*R + 1, 2v2 team_vs_bot, D=528
N_AGENTS = 2
WORKLOAD = "2v2 team_vs_bot_configs (2 agents vs 2 on-device smart bots), octagon map, auto-reset, random actions"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=None)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--envs-per-gpu", type=int, default=65536)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--ref-envs", type=int, default=8192, help="envs per step of the CPU reference arm (bounded sample)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--burn-in", type=int, default=600,
                   help="untimed ticks before the warm-up steps: episodes (mean ~190 ticks) desynchronise, so the timed "
                        "steps see the steady-state mix of episode phases instead of 65536 freshly reset battles")
    return p.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the step kernel from the committed ncu summary, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        if nv:
            for nm in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
                v = getattr(nv, "nvmlClocksThrottleReason" + nm, None)
                if v is not None:
                    names[v] = nm
        while not self._stop.is_set():
            try:
                if nv:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                else:
                    import subprocess
                    out = subprocess.check_output(
                        ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], timeout=5).decode().strip().split(",")
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    for nm, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def kernel_split(E, D, step_ms, sim_ms, peak):
    """Per-kernel view of the step.  SURVEY 8d's 7039 B/env-step = 2 S_state + 3 A + 4 R + 1 (simulation kernel: state
    in and out, actions, rewards, done) + 4 R D (rows kernel: the observation rows)."""
    sim_b = (ALG_BYTES_PER_ENV_STEP - 4 * N_AGENTS * D) * E
    rows_b = 4 * N_AGENTS * D * E
    rows_ms = max(step_ms - sim_ms, 1e-6)
    return {
        "env_kernel<true,false>": {"ms": sim_ms, "alg_bytes": sim_b, "achieved_gbs": sim_b / sim_ms / 1e6,
                                   "frac": sim_b / sim_ms / 1e6 / peak, "share_of_step": sim_ms / step_ms,
                                   "note": "timed alone (at_step without an observation buffer); latency-bound "
                                           "fp64 / integer simulation, see DESIGN.md"},
        "rows_kernel": {"ms": rows_ms, "alg_bytes": rows_b, "achieved_gbs": rows_b / rows_ms / 1e6,
                        "frac": rows_b / rows_ms / 1e6 / peak, "share_of_step": rows_ms / step_ms,
                        "note": "step minus simulation (the two overlap slightly: programmatic dependent launch)"},
    }


def cpu_arm(n_envs, steps, warmup, seed, threads, burn_in=0):
    """Times the oracle port (oracle/tank_oracle.c) on host threads: the reference's algorithm on the CPU."""
    import numpy as np
    from oracle import oracle as O
    from oracle.atrng import synthetic_actions
    from alphatank_b200.configs.config_teams import team_vs_bot_configs
    cfg = O.config_from_team_dict(team_vs_bot_configs, auto_reset=True)
    env = O.OracleEnv(cfg, n_envs, seed)
    env.reset()
    ids = np.arange(n_envs)
    for t in range(burn_in):                   # same steady-state protocol as the GPU arm (untimed)
        env.step(synthetic_actions(seed, ids, t, N_AGENTS), threads=threads)
    acts = [synthetic_actions(seed, ids, burn_in + t, N_AGENTS) for t in range(warmup + steps)]
    for t in range(warmup):
        env.step(acts[t], threads=threads)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        env.step(acts[t], threads=threads)
    dt = time.perf_counter() - t0
    return n_envs * N_AGENTS * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps if args.steps is not None else 20
    warmup = args.warmup if args.warmup is not None else 3
    threads = os.cpu_count() or 1
    value, dt = cpu_arm(args.ref_envs, steps, warmup, args.seed, threads, burn_in=args.burn_in)
    sample = "%d envs x %d ticks per run after %d untimed burn-in ticks (oracle C port of the reference env step, pthreads over envs)" % (
        args.ref_envs, steps, args.burn_in)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_step": args.ref_envs, "agents_per_env": N_AGENTS,
                   "note": "the reference itself is single-env interpreted Python needing pygame (not installable on "
                           "the box); this arm times its C restatement (oracle/) on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    steps = args.steps if args.steps is not None else 200
    warmup = max(args.warmup if args.warmup is not None else 10, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    ge.build_cuda()
    from alphatank_b200 import MultiAgentTeamEnv
    from alphatank_b200.configs.config_teams import team_vs_bot_configs

    E = args.envs_per_gpu
    env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=E, device=dev, seed=args.seed, env_id_base=rank * E,
                            map="octagon", terminate_time=None, auto_reset=True)
    core = env.core
    env.reset()
    # inputs resident in HBM before the timed region: a ring of pre-generated action tensors
    ring = min(warmup + steps, 128)
    acts = torch.empty((ring, E, N_AGENTS, 3), dtype=torch.int32, device=dev)
    for t in range(ring):
        core.synthetic_actions(t, acts[t])

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        stream = torch.cuda.current_stream(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for i in range(n):
            fn(i)
        ev1.record(stream)
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for t in range(args.burn_in):              # state preparation (like generating the synthetic inputs): not a step
        core.step(acts[t % ring])
    for t in range(warmup):
        core.step(acts[(args.burn_in + t) % ring])
    l0 = core.launch_count()
    with ClockSampler(local) as clk:
        ms = timed(lambda i: core.step(acts[(args.burn_in + warmup + i) % ring]), steps)
    launches = core.launch_count() - l0
    total_envs = E * world

    # per-kernel split of the step: the simulation kernel alone (at_step without an observation buffer launches only
    # env_kernel<true, false>); the rest of the step is rows_kernel
    import ctypes as C
    from alphatank_b200 import _capi
    sim_steps = max(min(steps, 100), 10)

    def sim_only(i):
        a = acts[(warmup + i) % ring]
        _capi.check(core.L.at_step(core.h, C.c_void_p(a.data_ptr()), None, C.c_void_p(core.rewards.data_ptr()),
                                   C.c_void_p(core.done.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "at_step")
    ms_sim = timed(sim_only, sim_steps) / sim_steps
    value = total_envs * N_AGENTS * steps / (ms * 1e-3)

    # end to end through the reference-facing API with HOST action / reward / done buffers
    h_acts = [acts[t % ring].cpu().pin_memory() for t in range(min(ring, 16))]
    e2e_steps = max(min(steps, 100), 10)
    for t in range(3):
        env.step_host(h_acts[t % len(h_acts)])
    ms_e2e = timed(lambda i: env.step_host(h_acts[i % len(h_acts)]), e2e_steps)
    e2e_value = total_envs * N_AGENTS * e2e_steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg_bytes = ALG_BYTES_PER_ENV_STEP * E
    kernel_s = ms * 1e-3 / steps
    achieved = alg_bytes / kernel_s / 1e9
    traffic = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": E, "total_envs": total_envs, "agents_per_env": N_AGENTS,
                   "burn_in_ticks": args.burn_in, "tanks_per_env": 4, "obs_dim": core.D, "parallelism": "env-id sharding x%d, no step-path collective"
                   % world, "timing": "per-step footprint (SoA state ~%d MB + %d MB of observations written) exceeds the "
                   "126 MB L2; no flush between steps needed" % (804 * E // 2**20, core.D * 4 * N_AGENTS * E // 2**20),
                   "env_steps_per_sec": value / N_AGENTS},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "kernel": "one step = env_kernel<true,false> (simulation) + rows_kernel (observation rows)",
                     "alg_bytes_per_launch": alg_bytes,
                     "alg_bytes_per_env_step": ALG_BYTES_PER_ENV_STEP, "peak_source": peak_src,
                     "kernel_ms": kernel_s * 1e3,
                     "kernels": kernel_split(E, core.D, kernel_s * 1e3, ms_sim, peak)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * N_AGENTS * 3 * 4,
                "d2h_bytes_per_step": E * N_AGENTS * 4 + E, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "note": "MultiAgentTeamEnv.step_host: pinned host actions in (DMA copy beside the previous step's rows "
                        "kernel), rewards+done posted into pinned host memory and read by the host every step (the "
                        "call returns when they have landed); observations stay in HBM for the policy"},
        "gpu_launches": launches * world,
        "clocks": clk.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_ref, t_ref = 4096, 8
        v, dt = cpu_arm(n_ref, t_ref, 2, args.seed, threads)
        while dt < 8.0 and t_ref < 4096:     # grow the sample to about 10-20 s of CPU work
            t_ref *= 2
            v, dt = cpu_arm(n_ref, t_ref, 2, args.seed, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d envs x %d ticks of the same workload, oracle C port, %.1f s" % (
                                    n_ref, t_ref, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
