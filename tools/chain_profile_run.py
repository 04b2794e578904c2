"""Small whole-chain run for ncu captures of the streaming kernels: python tools/chain_profile_run.py [frames]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine as E
from sdr_receiver_dvb_t2_b200.chain import FrameChain
from tools.modulator import Modulator

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tables = E.mode_tables(E.mode_init('32K', True, 7, '1/128', 59))
eng = t2.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
eng.set_option(E.OPT_DEMAP_SATURATE, 1)
mod = Modulator(tables, mod=3, cod=2, fec_normal=True, n_blocks=202, ti_len=3, seed=5)
fr = np.stack([mod.frame(noise_cn_db=20.5)['time'] for _ in range(2)])
x = torch.from_numpy(fr).cuda()[torch.arange(F) % 2].contiguous()
chain = FrameChain(eng, tables, mod=3, cod=2, fec_type=1, n_blocks=202, ti_len=3)
for _ in range(2):
    r = chain.decode_frames_fused(x)        # the product path: one t2b200_frames_decode call
torch.cuda.synchronize()
print('ok', float((r['trials_left'] >= 0).float().mean()))
