#!/usr/bin/env python
"""Debug aid (GPU box): run a sweep config on the CUDA path and on the CPU oracle side by side and, at the first tick
where anything differs, print both states of the first differing env (tanks, bullets, counters).
usage: python tools/diverge_probe.py CONFIG [N seed base ticks]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from gpu_env import GpuEnv  # noqa: E402
from oracle.atrng import synthetic_actions  # noqa: E402
from test_kernel_logic_cpu import SWEEP  # noqa: E402


def dump(tag, st, n):
    print("  [%s] n_bullets %d tick %d tstep %d epstep %d rng %d" % (tag, st["n_bullets"], st["tick"], st["training_step"],
                                                                  st["episode_steps"], st["rng_ctr"]))
    for i in range(n):
        t = st["tanks"][i]
        print("    tank %d x %.17g y %.17g a %d sp %d alive %d hw %d rew %.17g hit %d/%d shot %d" % (
            i, t["x"], t["y"], t["angle"], t["speed"], t["alive"], t["hitting_wall"], t["reward"], t["num_hit"],
            t["num_be_hit"], t["last_shot_ms"]))
    for s in range(st["n_bullets"]):
        b = st["bullets"][s]
        print("    bullet %d owner %d x %.17g y %.17g d (%.6f %.6f) dist %d bounces %d pend %.2f cleared %d" % (
            s, b["owner"], b["x"], b["y"], b["dx"], b["dy"], b["distance"], b["bounces"], b["pending"], b["cleared"]))


def main():
    name = sys.argv[1]
    N, seed, base, ticks = (int(x) for x in (sys.argv[2:6] + ["1003", "31", "900", "120"][len(sys.argv) - 2:]))
    spec = SWEEP[name]
    a = parity.make_oracle_env(spec, N, seed, base, auto_reset=True)
    b = GpuEnv(spec, N, seed, base, auto_reset=True)
    assert np.array_equal(a.reset(), b.reset()), "reset obs differ"
    ids = np.arange(base, base + N)
    prev = None
    for t in range(ticks):
        acts = synthetic_actions(seed, ids, t, a.A) if a.A else None
        oa, ra, da = a.step(acts, threads=os.cpu_count() or 1)
        ob, rb, db = b.step(acts)
        bad = np.unique(np.concatenate([np.nonzero(da != db)[0], np.nonzero((ra != rb).any(axis=1))[0],
                                        np.nonzero((oa != ob).any(axis=1))[0]]))
        if len(bad):
            e = int(bad[0])
            print("%s: first difference at tick %d, envs %s" % (name, t, bad[:10].tolist()))
            print("  obs cols", np.nonzero(oa[e] != ob[e])[0][:40].tolist())
            print("  rewards oracle %s cuda %s  done %d / %d" % (ra[e], rb[e], da[e], db[e]))
            if acts is not None:
                print("  actions", acts[e].tolist())
            dump("oracle", a.get_state(e), a.n_tanks)
            dump("cuda", b.get_state(e), a.n_tanks)
            if prev is not None:
                print("  -- state of env %d after tick %d (identical obs on both sides)" % (e, t - 1))
                dump("oracle t-1", prev[e], a.n_tanks)
            return 1
        prev = a.get_state()
    print("%s: %d envs x %d ticks identical" % (name, N, ticks))
    return 0


if __name__ == "__main__":
    sys.exit(main())
