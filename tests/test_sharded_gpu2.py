"""Multi-GPU FEC through the C-ABI (t2b200_comm_init / t2b200_ldpc_decode_sharded, csrc/comm.cu): run under torchrun with 2+
ranks on 2+ GPUs (`gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... tests/test_sharded_gpu2.py`);
not collected by pytest (needs several GPUs).  Rank 0 holds noisy codewords, every rank decodes a shard, the gathered bits
must equal a single-GPU decode of the same batch, for byte-per-bit and packed output and a ragged batch size."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
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
    ok = True
    for code, n in ((2, 4096 + 7 * 32), (7, 2500), (2, 40)):
        N, K, KB = eng.ldpc_geometry(code)
        llr = out = want = None
        if rank == 0:
            base, info = O.make_llr(code, 128, {2: 2.9, 7: 2.9}[code], seed=5)
            llr = torch.from_numpy(np.tile(base, ((n + 127) // 128, 1))[:n].copy()).to(dev)
            want = eng.ldpc_decode(code, llr, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE)['bits']
        for flags in (E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE, E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE | E.LDPC_PACK_BITS):
            if rank == 0:
                out = torch.zeros((n, KB // 8 if flags & E.LDPC_PACK_BITS else KB), dtype=torch.uint8, device=dev)
            eng.ldpc_decode_sharded(code, 0, llr, n, out=out, flags=flags)
            eng.sync()
            if rank == 0:
                got = out.cpu().numpy()
                if flags & E.LDPC_PACK_BITS:
                    got = np.unpackbits(got, axis=1)
                same = np.array_equal(got, want.cpu().numpy())
                print('code %d n %d flags %d: %s' % (code, n, flags, 'ok' if same else 'MISMATCH'), flush=True)
                ok &= same
    eng.comm_destroy()
    eng.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print('SHARDED_OK' if flag.item() else 'SHARDED_FAIL', flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
