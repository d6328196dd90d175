"""Multi-GPU plumbing of the path: clips (videos) are independent units, so ranks only ever (a) split the work list
and (b) agree on the slowest rank's time.  No data-path collective."""
import torch


def shard(items, rank, world):
    """Round-robin split of independent clips over ranks (reference: InferenceSampler over videos,
    projects/IDOL/idol/data/build.py:234-247)."""
    return list(items)[rank::world]


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (device-timed intervals are reported as the slowest rank's)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
