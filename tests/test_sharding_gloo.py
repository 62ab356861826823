"""CPU, world_size 2 over gloo: the multi-GPU path of the step is pure env-id sharding.  Each rank steps its own
contiguous id range (here with the kernel-logic emulator standing in for its GPU); the gathered result equals
one process stepping the whole id space, and the timing reduction is the max over ranks."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TOTAL, TICKS, SEED = 48, 60, 31


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (REPO, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emul_env
    from alphatank_b200 import sharding
    from oracle.atrng import synthetic_actions
    from test_kernel_logic_cpu import SWEEP
    first, count = sharding.shard_range(TOTAL, world, rank)
    env = emul_env.make_emul_env(SWEEP["team_vs_bot"], count, SEED, first, auto_reset=True)
    env.reset()
    ids = np.arange(first, first + count)
    obs_log, rew_log, done_log = [], [], []
    for t in range(TICKS):
        o, r, d = env.step(synthetic_actions(SEED, ids, t, env.A))     # no collective on the step path
        obs_log.append(o.copy()); rew_log.append(r.copy()); done_log.append(d.copy())
    dist.barrier()
    mx = sharding.max_over_ranks(10.0 * (rank + 1))
    assert mx == 10.0 * world
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), obs=np.stack(obs_log), rew=np.stack(rew_log),
             done=np.stack(done_log), first=first, count=count)
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import emul_env
    from oracle.atrng import synthetic_actions
    from test_kernel_logic_cpu import SWEEP
    env = emul_env.make_emul_env(SWEEP["team_vs_bot"], TOTAL, SEED, 0, auto_reset=True)
    env.reset()
    ids = np.arange(TOTAL)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["first"]) for p in parts] == [0, 24] and [int(p["count"]) for p in parts] == [24, 24]
    for t in range(TICKS):
        o, r, d = env.step(synthetic_actions(SEED, ids, t, env.A))
        assert np.array_equal(o, np.concatenate([p["obs"][t] for p in parts]))
        assert np.array_equal(r, np.concatenate([p["rew"][t] for p in parts]))
        assert np.array_equal(d, np.concatenate([p["done"][t] for p in parts]))
