#!/usr/bin/env python
"""Key metrics of .ncu-rep files (ncu -i ... --page raw --csv) as one table; with --traffic-json also writes
profiles/step_kernel_traffic.json (DRAM bytes of the step's two kernels, stamped with the commit they were captured at).
usage: tools/ncu_summary.py [--traffic-json COMMIT SIM.ncu-rep ROWS.ncu-rep] REP..."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def num(d, k):
    v, u = d.get(k, ("nan", ""))
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return float("nan"), u
    return x, u


def to_bytes(x, u):
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


args = sys.argv[1:]
traffic = None
if args and args[0] == "--traffic-json":
    traffic, args = args[1:4], args[4:]
reps = ([traffic[1], traffic[2]] if traffic else []) + args
data = {r: read(r) for r in reps}
names = [data[r].get("Kernel Name", ("?", ""))[0][:44] for r in reps]
print("%-66s" % "" + "  ".join("%-22s" % os.path.basename(r)[:22] for r in reps))
print("%-66s" % "kernel" + "  ".join("%-22s" % n[:22] for n in names))
for k in KEYS:
    cells = []
    for r in reps:
        x, u = num(data[r], k)
        cells.append("%-22s" % ("%.4g %s" % (x, u)))
    print("%-66s" % k + "  ".join(cells))
if traffic:
    commit, sim, rows = traffic
    rec = {"kernel": "one step = env_kernel (simulation) + rows_kernel (observation rows)", "captured_at_commit": commit,
           "capture": "ncu --set full --clock-control none at steady state (bench.py --burn-in 300, launch 322 / 320 of the kernel): %s, %s" % (
               os.path.basename(sim), os.path.basename(rows)), "envs": 65536, "per_kernel": {}}
    tot_r = tot_w = 0.0
    for rep in (sim, rows):
        d = data[rep]
        r = to_bytes(*num(d, "dram__bytes_read.sum"))
        w = to_bytes(*num(d, "dram__bytes_write.sum"))
        t, tu = num(d, "gpu__time_duration.sum")
        t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(tu, 1)
        rec["per_kernel"][d["Kernel Name"][0]] = {"dram_bytes_read": int(r), "dram_bytes_write": int(w), "gpu_time_us_under_ncu": t_us}
        tot_r += r
        tot_w += w
    rec["kernels"] = list(rec["per_kernel"])
    rec["dram_bytes_read"], rec["dram_bytes_write"], rec["dram_bytes_per_launch"] = int(tot_r), int(tot_w), int(tot_r + tot_w)
    json.dump(rec, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "step_kernel_traffic.json"), "w"), indent=1)
    print("wrote profiles/step_kernel_traffic.json:", rec["dram_bytes_per_launch"], "bytes per step")
