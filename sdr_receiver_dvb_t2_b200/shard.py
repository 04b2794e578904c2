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


class CodewordSharder:
    """SURVEY 8e scatter / gather variant: ONE rank demodulates (it holds the int8 LLRs of a pooled batch), every
    rank decodes a contiguous shard of whole 32-codeword lock-step groups, the BBFRAME bits come back to the
    demodulating rank in codeword order.  The only data-path traffic is one point-to-point transfer each way per
    rank (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests): 64 800 B out and k_out B back per codeword.

    decode_fn(llr[n][n_bits] int8) -> bits[n][k_out] uint8 on the same device (the rank's own t2b200 engine).
    """

    def __init__(self, decode_fn, n_bits, k_out, src=0, device='cpu', granule=32):
        self.decode_fn, self.n_bits, self.k_out, self.src, self.device, self.granule = decode_fn, n_bits, k_out, src, device, granule
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def spans(self, n_cw):
        return [shard_range(n_cw, r, self.world, self.granule) for r in range(self.world)]

    def decode(self, llr, n_cw, out=None):
        """llr: int8[n_cw][n_bits] on the src rank (ignored elsewhere).  Returns bits uint8[n_cw][k_out] on src, None
        on the other ranks.  Every rank must call it with the same n_cw."""
        spans = self.spans(n_cw)
        lo, hi = spans[self.rank]
        if self.world == 1:
            return self.decode_fn(llr)
        if self.rank == self.src:
            ops = [dist.P2POp(dist.isend, llr[a:b], r) for r, (a, b) in enumerate(spans) if r != self.src and b > a]
            mine = llr[lo:hi]
        else:
            mine = torch.empty((hi - lo, self.n_bits), dtype=torch.int8, device=self.device)
            ops = [dist.P2POp(dist.irecv, mine, self.src)] if hi > lo else []
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        bits = self.decode_fn(mine) if hi > lo else torch.empty((0, self.k_out), dtype=torch.uint8, device=self.device)
        if self.rank == self.src:
            if out is None:
                out = torch.empty((n_cw, self.k_out), dtype=torch.uint8, device=self.device)
            out[lo:hi] = bits
            ops = [dist.P2POp(dist.irecv, out[a:b], r) for r, (a, b) in enumerate(spans) if r != self.src and b > a]
        else:
            ops = [dist.P2POp(dist.isend, bits, self.src)] if hi > lo else []
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out if self.rank == self.src else None
