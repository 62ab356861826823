#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture per source function / line.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv
    python tools/ncu_lines.py X.csv [--top 40] [--envs 65536]

Per source line the csv carries: warp-level instructions executed, thread-level instructions, stall samples by
reason.  Lines are attributed to the enclosing function of at_sim.h / at_kernels.cu (brace-depth scan of
`AT_HD ... name(` definitions).  Development tooling only (profiles/*.md are produced from its output)."""
import argparse
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def func_ranges(path):
    """(start_line, end_line, name) of function bodies, found with a light brace scan."""
    out, stack = [], []
    depth = 0
    pat = re.compile(r"\b([A-Za-z_][A-Za-z_0-9]*)\s*\([^;{}]*\)\s*(?:const)?\s*(?::[^{;]*)?\{\s*(?://.*)?$")
    kw = {"if", "for", "while", "switch", "else", "do", "return"}
    pending = None
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return out
    buf = ""
    for ln, s in enumerate(lines, 1):
        code = s.split("//")[0]
        buf = (buf + " " + code.strip()) if buf and not buf.rstrip().endswith((";", "}", "{")) else code.strip()
        m = pat.search(buf)
        if m and m.group(1) not in kw and "=" not in buf.split(m.group(1))[0][-3:]:
            pending = m.group(1)
        for ch in code:
            if ch == "{":
                depth += 1
                if pending is not None:
                    stack.append((pending, ln, depth))
                    pending = None
            elif ch == "}":
                if stack and stack[-1][2] == depth:
                    nm, st, _ = stack.pop()
                    out.append((st, ln, nm))
                depth -= 1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--by", default="samples", choices=["samples", "inst"], help="order of the per-line list")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    cur_file, hdr = None, None
    per_line = {}
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        stalls = {k[6:]: num(k) for k in d if k.startswith("stall_") and "(Not Issued)" not in k}
        per_line[(cur_file, int(r[0]))] = dict(
            src=r[1], inst=num("Instructions Executed"), tinst=num("Thread Instructions Executed"),
            samples=num("# Samples"), stalls=stalls)
    files = sorted({f for f, _ in per_line})
    tot_inst = sum(v["inst"] for v in per_line.values())
    tot_t = sum(v["tinst"] for v in per_line.values())
    tot_s = sum(v["samples"] for v in per_line.values())
    warps = a.envs / 32
    print("total warp-instr %.3e (%.0f / warp), thread-instr %.3e (%.0f / env-step), lanes %.2f, samples %d" % (
        tot_inst, tot_inst / warps, tot_t, tot_t / a.envs, tot_t / max(tot_inst, 1), tot_s))
    allst = defaultdict(float)
    for v in per_line.values():
        for k, x in v["stalls"].items():
            allst[k] += x
    print("stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * x / max(tot_s, 1)) for k, x in
                                        sorted(allst.items(), key=lambda kv: -kv[1])[:9]))
    byfn = defaultdict(lambda: dict(inst=0.0, tinst=0.0, samples=0.0, stalls=defaultdict(float)))
    for f in files:
        rng = func_ranges(f if os.path.exists(f) else os.path.join(ROOT, "alphatank_b200", "csrc", os.path.basename(f)))
        for (ff, ln), v in per_line.items():
            if ff != f:
                continue
            name = os.path.basename(f)
            best = None
            for st, en, nm in rng:
                if st <= ln <= en and (best is None or st >= best[0]):
                    best = (st, nm)
            if best:
                name = best[1]
            b = byfn[name]
            b["inst"] += v["inst"]; b["tinst"] += v["tinst"]; b["samples"] += v["samples"]
            for k, x in v["stalls"].items():
                b["stalls"][k] += x
    print("\n%-22s %7s %9s %6s %8s  top stalls" % ("function", "warp%", "inst/warp", "lanes", "sample%"))
    for nm, b in sorted(byfn.items(), key=lambda kv: -kv[1]["samples"])[:a.top]:
        ts = sorted(b["stalls"].items(), key=lambda kv: -kv[1])[:4]
        print("%-22s %6.1f%% %9.0f %6.1f %7.1f%%  %s" % (
            nm, 100 * b["inst"] / tot_inst, b["inst"] / warps, b["tinst"] / max(b["inst"], 1),
            100 * b["samples"] / max(tot_s, 1),
            " ".join("%s:%.0f%%" % (k, 100 * x / max(b["samples"], 1)) for k, x in ts)))
    print("\ntop lines by " + ("stall samples" if a.by == "samples" else "warp instructions"))
    for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][a.by])[:a.top]:
        ts = sorted(v["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("%5.2f%% inst %5.2f%% lanes %4.1f %s:%d  %s  [%s]" % (
            100 * v["samples"] / max(tot_s, 1), 100 * v["inst"] / tot_inst, v["tinst"] / max(v["inst"], 1),
            os.path.basename(f), ln, v["src"].strip()[:90],
            " ".join("%s:%.0f%%" % (k, 100 * x / max(v["samples"], 1)) for k, x in ts)))


if __name__ == "__main__":
    sys.exit(main())
