"""N > 1 host logic on CPU: world_size-2 gloo process group -- shard ranges cover the batch exactly once on
32-codeword boundaries, the job time is the max over ranks, counts gather in rank order."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdr_receiver_dvb_t2_b200.shard import CodewordSharder, gather_counts, max_over_ranks, shard_range, sum_over_ranks


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(4040, rank, world, granule=32)
    t = max_over_ranks(10.0 + rank)
    total = sum_over_ranks(hi - lo)
    counts = gather_counts(hi - lo)
    q.put((rank, lo, hi, t, total, counts))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, tot0, c0), (r1, lo1, hi1, t1, tot1, c1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == 4040            # disjoint, complete
    assert lo1 % 32 == 0                                      # lock-step groups stay whole
    assert t0 == t1 == 11.0                                   # max over ranks
    assert tot0 == tot1 == 4040 and c0 == c1 == [hi0 - lo0, hi1 - lo1]


def test_shard_range_properties():
    for n in (0, 1, 31, 32, 33, 4040, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world, 32) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0] and a[1] % 32 == 0 or a[1] == n


def _hard(llr):
    """stand-in decoder for the plumbing test: hard decision of the first 40 LLRs, tagged with the decoding rank"""
    b = (llr[:, :40] < 0).to(torch.uint8)
    b[:, 39] = dist.get_rank()
    return b


def _sg_worker(rank, world, port, q, n_cw):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    llr = torch.randint(-128, 128, (n_cw, 64), generator=g, dtype=torch.int8) if rank == 0 else None
    sh = CodewordSharder(_hard, 64, 40, src=0)
    out = sh.decode(llr, n_cw)
    if rank == 0:
        want = (llr[:, :40] < 0).to(torch.uint8)
        ok = bool((out[:, :39] == want[:, :39]).all())
        owners = out[:, 39].tolist()
        q.put((ok, owners, sh.spans(n_cw)))
    else:
        assert out is None
    dist.destroy_process_group()


def test_codeword_scatter_gather_two_ranks():
    """rank 0 scatters LLR shards (whole 32-groups, ragged tail), both ranks 'decode', bits return in order"""
    world, n_cw = 2, 32 * 3 + 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_sg_worker, args=(r, world, port, q, n_cw)) for r in range(world)]
    for p in ps:
        p.start()
    ok, owners, spans = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert spans[0][1] % 32 == 0 and spans[1][1] == n_cw
    for r, (a, b) in enumerate(spans):
        assert all(o == r for o in owners[a:b])             # each shard was decoded by its rank
