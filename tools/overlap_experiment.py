"""Do the streaming kernels run underneath a resident LDPC decoder?  (development aid)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine as E
from oracle import pyoracle as O

hi = int(os.environ.get('HI', '0'))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream(priority=-1 if hi else 0)
ea, eb = t2.Engine(0, stream=sa.cuda_stream), t2.Engine(0, stream=sb.cuda_stream)
base, _ = O.make_llr(2, 256, 2.9, seed=2)
llr = torch.from_numpy(np.tile(base, (16, 1))[:4032].copy()).cuda()
out = torch.empty((4032, 43040), dtype=torch.uint8, device='cuda')
x = torch.randn((1200, 32768, 2), device='cuda').view(torch.float32)
x = torch.view_as_complex(x.reshape(1200, 32768, 2).contiguous())
y = torch.empty_like(x)
flags = (0 if os.environ.get('NATIVE') else E.LDPC_GROUP32) | E.LDPC_BCH_DESCRAMBLE


def ldpc():
    with torch.cuda.stream(sa):
        ea.ldpc_decode(2, llr, flags=flags, out=out, want_status=False)


def fft(n):
    with torch.cuda.stream(sb):
        for _ in range(n):
            eb.fft(x, out=y)


def timed(f):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    f()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


for _ in range(2):
    ldpc(); fft(2)
print('ldpc alone %.2f ms' % timed(ldpc))
print('fft x10 alone %.2f ms' % timed(lambda: fft(10)))
print('both %.2f ms' % timed(lambda: (ldpc(), fft(10))))
print('both (fft first) %.2f ms' % timed(lambda: (fft(10), ldpc())))
