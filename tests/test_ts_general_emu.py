"""N1, normal mode and mixed-mode batches: the frame plan of csrc/ts_general.h (copy segments + CRC tasks, the same source
the GPU kernel runs) executed on the CPU (tests/cpp/ts_emu.cpp) against the oracle port of the reference's byte-serial
bb_de_header (oracle/port/ts_port.c, pinned to the compiled reference and its golden datagrams by tests/test_oracle_ts.py):
datagrams byte for byte, for HEM, NM, mode flips by header bit errors, corrupted CRC bytes, lost frames and every SYNCD
resynchronisation branch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.test_oracle_ts import CASES, make
from tests.ts_helpers import bbframes, header

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def emu():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'tests', 'cpp', 'libts_emu.so')
        src = os.path.join(ROOT, 'tests', 'cpp', 'ts_emu.cpp')
        deps = [src, os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'csrc', 'ts_general.h')]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-w', '-o', so, src], check=True)
        _lib = C.CDLL(so)
        _lib.emu_ts_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return _lib


def emu_datagrams(frames):
    """-> datagrams; stops in front of a high-efficiency-mode frame entered with a packet index beyond 188 (normal mode's
    too-short-SYNCD resynchronisation leaves that behind): the reference then skips -(188 - index) bytes FORWARD
    (bb_de_header.cpp:360-364) and reads arbitrarily far outside its input -- undefined there, zeros on the GPU"""
    L = emu()
    st = np.zeros(L.emu_ts_state_size(), np.uint8)
    out = []
    for f in frames:
        f = np.ascontiguousarray(f, np.uint8)
        reg = 0
        for b in f[:80]:
            x = (int(b) ^ reg) & 1
            reg >>= 1
            if x:
                reg ^= 0xAB
        split, idx_packet = st.view(np.int32)[:2]
        if reg == 0xAB and split and idx_packet > 188:
            break
        buf = np.zeros(len(f) // 8 + 2 * 188 + 64, np.uint8)
        n = L.emu_ts_frame(st.ctypes.data, f.ctypes.data, len(f), buf.ctypes.data, len(buf))
        assert n > -9, (n, st.view(np.int32)[:4])
        out.append(None if n < 0 else buf[:n].copy())
    return out


def port_datagrams(frames):
    p = O.PortTs()
    return [p.feed(f) for f in frames]


def same(a, b):
    b = b[:len(a)]
    for i, (x, y) in enumerate(zip(a, b)):
        assert (x is None) == (y is None), i
        if x is not None:
            assert len(x) == len(y) and np.array_equal(x, y), (i, len(x), len(y), np.flatnonzero(x[:min(len(x), len(y))] != y[:min(len(x), len(y))])[:8])


@pytest.mark.parametrize('case', list(CASES))
def test_plan_equals_the_oracle_on_the_golden_cases(case):
    frames, _ = make(case)
    same(emu_datagrams(frames), port_datagrams(frames))


def mixed_stream(seed):
    """HEM and NM stretches back to back, ragged data fields, header faults of every kind, corrupted CRC bytes on air"""
    rng = np.random.default_rng(seed)
    k_bch = int(rng.choice([9552, 43040, 7032, 53840]))
    cap = (k_bch - 80) // 8 - 40                       # room for normal mode's reads behind the data field
    parts = []
    for _ in range(int(rng.integers(2, 5))):
        hem = bool(rng.integers(0, 2))
        n = int(rng.integers(2, 7))
        dfl = [int(rng.integers(1, cap)) if rng.random() < 0.4 else cap - int(rng.integers(0, 3)) for _ in range(n)]
        kinds = ['syncd65535', 'syncd_plus', 'syncd_minus', 'crc']
        faults = {int(rng.integers(0, n)): kinds[int(rng.integers(0, 4))] for _ in range(int(rng.integers(0, 3)))}      # one per frame
        faults = tuple(faults.items())
        fr, _ = bbframes(k_bch, dfl, n, hem, rng, faults)
        for f in fr:
            if rng.random() < 0.3:                     # a payload bit error: some packet's CRC no longer matches
                f[80 + int(rng.integers(0, 8 * min(dfl)))] ^= 1
        parts.append(fr)
    return np.concatenate(parts)


@pytest.mark.parametrize('seed', range(40))
def test_plan_equals_the_oracle_on_mixed_streams(seed):
    frames = mixed_stream(seed)
    got = emu_datagrams(frames)
    assert len(got) >= 2
    same(got, port_datagrams(frames))


def test_oversize_data_field_is_dropped():
    frames, _ = bbframes(9552, 1100, 4, False, np.random.default_rng(1))
    frames[2, :80] = header(9552, 0, False)
    got = emu_datagrams(frames)
    assert got[2] is None and got[3] is not None
