"""Tiny LDPC run for ncu captures: python tools/ldpc_profile_run.py <code> <n_cw> <flags> [Eb/N0 dB]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine as E
from oracle import pyoracle as O
code = int(sys.argv[1]); n = int(sys.argv[2]); flags = int(sys.argv[3])
eng = t2.Engine(0)
eb = float(sys.argv[4]) if len(sys.argv) > 4 else 2.9
base, info = O.make_llr(code, 32, eb, seed=2)
llr = torch.from_numpy(np.tile(base, ((n + 31) // 32, 1))[:n].copy()).cuda()
for _ in range(2):
    r = eng.ldpc_decode(code, llr, flags=flags)
eng.sync()
print('iters', r['iterations'].float().mean().item())
