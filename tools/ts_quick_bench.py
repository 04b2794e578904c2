"""Quick timing of t2b200_ts_packetize on 4040 normal-FECFRAME BBFRAMEs (development aid)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import sdr_receiver_dvb_t2_b200 as t2
from tests.ts_helpers import bbframes

eng = t2.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
fr, _ = bbframes(43040, 5370, 64, True, np.random.default_rng(1))
x = torch.from_numpy(np.tile(fr, (64, 1))[:4040].copy()).cuda()
for _ in range(2):
    eng.ts_packetize(x)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.ts_packetize(x)
b.record(); torch.cuda.synchronize()
print('ts_packetize 4040 frames: %.3f ms' % (a.elapsed_time(b) / 5))
