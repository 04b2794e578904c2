"""Timing of the sharded FEC stage alone (development aid): torchrun --nproc-per-node N tools/sharded_bench.py [codewords per rank]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import torch.distributed as dist
import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine as E
from oracle import pyoracle as O

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
eng = t2.Engine(local, stream=torch.cuda.current_stream().cuda_stream)
ident = [E.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ident, src=0)
eng.comm_init(rank, world, ident[0])
per = int(sys.argv[1]) if len(sys.argv) > 1 else 4032
code = 2
N, K, KB = eng.ldpc_geometry(code)
flags = E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE
llr = out = None
if rank == 0:
    # lock-step groups of different difficulty (like demapped LLRs: the iteration count of a group is its slowest codeword's)
    rng = np.random.default_rng(1)
    pools = [O.make_llr(code, 32, eb, seed=2 + k)[0] for k, eb in enumerate((2.5, 2.7, 2.9, 2.9, 3.1, 3.3))]
    pick = rng.integers(0, len(pools), (per + 31) // 32)
    shard = np.concatenate([pools[k] for k in pick])[:per]
    llr = torch.from_numpy(np.tile(shard, (world, 1))).to(dev)
    out = torch.empty((per * world, KB), dtype=torch.uint8, device=dev)
    one = torch.empty((per, KB), dtype=torch.uint8, device=dev)
    eng.ldpc_decode(code, llr[:per], flags=flags, out=one, want_status=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.ldpc_decode(code, llr[:per], flags=flags, out=one, want_status=False)
    e1.record(); torch.cuda.synchronize()
    single = e0.elapsed_time(e1) / 3
for it in range(5):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.ldpc_decode_sharded(code, 0, llr, per * world, out=out, flags=flags)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1)
        print('sharded %d x %d cw: %.2f ms -> %.0f cw/s (one GPU, %d cw: %.2f ms; efficiency %.2f)' % (world, per, ms, per * world / ms * 1e3, per, single, single / ms), flush=True)
eng.comm_destroy(); eng.close()
dist.destroy_process_group()
