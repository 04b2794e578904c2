"""The check-node arithmetic of the CUDA LDPC decoder (csrc/ldpc_pair.h: two codewords per register in s16x2 halves, DPX /
PRMT instructions restated in C for the host) executed on the CPU in the kernel's own order -- thread by thread, barrier
phase by barrier phase, level by level where two check nodes of a layer share a bit (tests/cpp/ldpc_pair_emu.cpp) -- against
the oracle port of the reference decoder: trial counts and all int8 posteriors of both codewords must be identical for all
15 codes, for converging, non-converging and saturating inputs.  No GPU needed; the GPU tests (test_ldpc_gpu.py) check the
kernel itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def emu():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'tests', 'cpp', 'libldpc_pair_emu.so')
        src = [os.path.join(ROOT, 'tests', 'cpp', 'ldpc_pair_emu.cpp'),
               os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'csrc', 'ldpc_schedule.cpp')]
        deps = src + [os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'csrc', 'ldpc_pair.h')]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-w', '-o', so] + src, check=True)
        _lib = C.CDLL(so)
        _lib.emu_ldpc_decode_pair.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    return _lib


def emu_decode(code, llr, trials):
    llr = np.ascontiguousarray(llr, np.int8)
    post = np.empty_like(llr)
    it = C.c_int()
    r = emu().emu_ldpc_decode_pair(code, llr.ctypes.data, trials, post.ctypes.data, C.byref(it))
    return r, post, it.value


EBN0 = {0: 1.3, 1: 2.4, 2: 2.9, 3: 3.5, 4: 3.9, 5: 4.4, 6: 1.2, 7: 2.6, 8: 3.0, 9: 3.6, 10: 4.1, 11: 4.8, 12: 0.5, 13: 0.8, 14: 1.0}


@pytest.mark.parametrize('code', list(range(15)))
def test_pair_arithmetic_equals_the_oracle(code):
    llr, _ = O.make_llr(code, 2, EBN0[code], seed=100 + code)
    tr, _, post = O.port_ldpc_decode(code, llr, 25, want_post=True)
    r, p, it = emu_decode(code, llr, 25)
    assert r == tr and it == 25 - tr if tr >= 0 else r == tr
    assert np.array_equal(p, post)


@pytest.mark.parametrize('code', [0, 2, 3, 5, 7, 11])
def test_pair_arithmetic_nonconverging_and_saturating(code):
    rng = np.random.default_rng(code)
    llr, _ = O.make_llr(code, 2, EBN0[code] - 1.5, seed=7 + code)              # hopeless: runs out of trials
    tr, _, post = O.port_ldpc_decode(code, llr, 6, want_post=True)
    r, p, _ = emu_decode(code, llr, 6)
    assert r == tr == -1 and np.array_equal(p, post)
    big = llr.astype(np.int32) * 6                                             # mostly +-127 / -128: the int8 saturation paths
    big = np.clip(big, -128, 127).astype(np.int8)
    big[rng.random(big.shape) < 0.02] = 0                                       # and exact zeros (they fail the parity test)
    tr, _, post = O.port_ldpc_decode(code, big, 8, want_post=True)
    r, p, _ = emu_decode(code, big, 8)
    assert r == tr and np.array_equal(p, post)
