"""Multi-GPU plumbing of the env step: envs are independent, so a job of ``total_envs`` is cut into contiguous
env-id ranges, one per rank / GPU, and NOTHING is exchanged on the step path.  Each env draws from the ATRNG
stream keyed by its GLOBAL id, so its trajectory does not depend on how the id space is cut (tests check this).
torch.distributed is only used to line the ranks up for timing (barrier + max over ranks) and, in a learner, for
the PPO gradient all-reduce."""
import os


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total_envs, world_size, rank):
    """Contiguous [first, first + count) of global env ids owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
