"""Quick LDPC-only timing (development aid; bench.py is the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine as E
from oracle import pyoracle as O

eng = t2.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
codes = [int(c) for c in sys.argv[1].split(',')] if len(sys.argv) > 1 else [2]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for code in codes:
    eb = {0: 1.4, 1: 2.5, 2: 2.9, 3: 3.4, 4: 3.9, 5: 4.3, 7: 2.9}.get(code, 3.0)
    base, info = O.make_llr(code, 256, eb, seed=2)
    llr = torch.from_numpy(np.tile(base, (nb // 256, 1))).cuda()
    N, K, KB = eng.ldpc_geometry(code)
    for flags, name in [(E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE, 'group32'), (E.LDPC_BCH_DESCRAMBLE, 'native')]:
        out = torch.empty((nb, KB), dtype=torch.uint8, device='cuda')
        r = eng.ldpc_decode(code, llr, flags=flags, out=out)
        torch.cuda.synchronize()
        it = r['iterations'].float().mean().item()
        ok = (r['trials_left'] >= 0).all().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.ldpc_decode(code, llr, flags=flags, out=out, want_status=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('code %d %s: %d cw in %.3f ms -> %.0f cw/s, mean iters %.2f, %.1f us/cw-iter/SM-slot, ok=%s'
              % (code, name, nb, ms, nb / ms * 1e3, it, ms * 1e3 / (nb * it / 148), ok), flush=True)
