"""Multi-GPU plumbing: FEC blocks / T2 frames are independent, so ranks work on disjoint contiguous shards and the
only cross-rank traffic is bookkeeping (counts, max-over-ranks timing).  Backend-agnostic (NCCL on GPUs, gloo in the
CPU tests); no data-path collective."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world, granule=1):
    """contiguous [lo, hi) of n_items for `rank`; shard boundaries fall on multiples of `granule` (32 keeps the
    reference's lock-step LDPC groups whole)"""
    units = (n_items + granule - 1) // granule
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * granule, n_items), min(hi_u * granule, n_items)


def max_over_ranks(value, device='cpu'):
    """the slowest rank's time is the job's time"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device='cpu'):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_counts(count, device='cpu'):
    """every rank's item count, on every rank (to place shard outputs in a global order)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
