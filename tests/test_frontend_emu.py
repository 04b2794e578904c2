"""The front-end kernels (csrc/frontend_kernels.h: DC removal, IQ correction, NCO, Farrow resampler, half-band decimator in
closed form per chunk) executed on the CPU -- the same source the GPU compiles, every barrier phase as a plain loop
(tests/cpp/frontend_emu.cpp) -- against the oracle port of the reference's serial loops (oracle/port/frontend_port.c).
No GPU needed; tests/test_frontend_gpu.py checks the kernels themselves.

Tolerances: the NCO phase recurrence is reproduced exactly (it indexes a sin / cos table, so it has to be); the DC average,
the resampler phase and the IQ statistics are evaluated in another order than the reference's sample-by-sample float
recurrences, so samples agree to 2e-6 of the RMS and the statistics to 2e-5 relative."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'csrc')
_lib = None

STREAM = np.dtype([('dc_re', 'f4'), ('dc_im', 'f4'), ('frequency_nco', 'f4'), ('x1', 'f4'), ('delay', 'f4', (3, 2)),
                   ('hist', 'f4', (63, 2)), ('parity', 'i4'), ('pad', 'i4')])
CHUNK = np.dtype([('len_in', 'i4'), ('short_to_float', 'f4'), ('c1', 'f4'), ('c2', 'f4'), ('frequency_est_filtered', 'f4'),
                  ('phase_nco', 'f4'), ('resample', 'f4')])
RESULT = np.dtype([('len_out', 'i4'), ('len_interp', 'i4'), ('theta', 'f4', (3,))])


def emu():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'tests', 'cpp', 'libfrontend_emu.so')
        src = os.path.join(ROOT, 'tests', 'cpp', 'frontend_emu.cpp')
        deps = [src, os.path.join(CSRC, 'frontend_kernels.h'), os.path.join(CSRC, 'frontend_tables.h')]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-w', '-o', so, src],
                           check=True)
        _lib = C.CDLL(so)
        _lib.emu_nco_run.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        _lib.emu_nco_end.argtypes = [C.c_float, C.c_float, C.c_int]
        _lib.emu_nco_end.restype = C.c_float
        _lib.emu_nco_stress.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        _lib.emu_fe_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                      C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        _lib.emu_cp_correlate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.emu_p1_correlate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        assert _lib.emu_fe_state_size() == STREAM.itemsize
    return _lib


def serial_nco(v, c, n):
    """dvbt2_demodulator.cpp:194-200 in numpy float32 scalars"""
    two_pi = np.float32(np.float32(3.14159265358979323846) * np.float32(2.0))
    v, c = np.float32(v), np.float32(c)
    out = np.empty(n, np.float32)
    for i in range(n):
        v = np.float32(v + c)
        while v > two_pi:
            v = np.float32(v - two_pi)
        while v < -two_pi:
            v = np.float32(v + two_pi)
        out[i] = v
    return out


NCO_CASES = [(0.0, 0.0, 300), (0.0, -1.9444850e-07, 3000), (3.2850355e-03, -2.28e-07, 3000), (1.0, 1e-8, 500), (1.0, -2.0 ** -25, 500),
             (1.0, -2.0 ** -24, 500), (1.0, 2.0 ** -24, 500), (4.0, -3.0e-3, 5000), (-4.0, 3.0e-3, 5000), (6.28, 1.1e-3, 5000),
             (-6.28, -1.1e-3, 5000), (0.5, -2.0 ** -25 * 1.5, 700), (2.0, -1.7e-7, 900), (1e-30, 1e-38, 100), (0.3, 0.77, 200),
             (6.0, 3.1, 100), (2.0 - 2.0 ** -22, 2.0 ** -23 * 0.75, 64), (1.9999, 2.0 ** -24 * 0.5, 3000)]


@pytest.mark.parametrize('case', range(len(NCO_CASES)))
def test_nco_recurrence_in_jumps_is_bit_exact(case):
    v, c, n = NCO_CASES[case]
    want = serial_nco(v, c, n)
    got = np.empty(n, np.float32)
    ns = C.c_int()
    emu().emu_nco_run(v, c, n, got.ctypes.data, C.byref(ns))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (np.flatnonzero(got != want)[:5], ns.value)
    assert np.float32(emu().emu_nco_end(v, c, n)).view(np.uint32) == want[-1].view(np.uint32)


def test_nco_recurrence_random():
    rng = np.random.default_rng(11)
    for _ in range(300):
        v = np.float32(rng.uniform(-6.28, 6.28)) if rng.random() < 0.8 else np.float32(rng.uniform(-1, 1) * 10.0 ** rng.uniform(-8, 0))
        c = np.float32(rng.choice([-1, 1]) * 10.0 ** rng.uniform(-9, -1.5))
        n = int(rng.integers(1, 1025))
        want = serial_nco(v, c, n)
        got = np.empty(n, np.float32)
        ns = C.c_int()
        emu().emu_nco_run(v, c, n, got.ctypes.data, C.byref(ns))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (float(v), float(c), n, np.flatnonzero(got != want)[:5])
        assert np.float32(emu().emu_nco_end(v, c, n)).view(np.uint32) == want[-1].view(np.uint32)


def test_nco_recurrence_stress():
    """200 000 random (phase, decrement, length) triples, incl. exact ties and powers of two, against real float additions
    (compared in C, tests/cpp/frontend_emu.cpp::emu_nco_stress; 2 000 000 cases were run once when the planner was written)"""
    v, c, n = C.c_float(), C.c_float(), C.c_int()
    bad = emu().emu_nco_stress(12345, 200000, C.byref(v), C.byref(c), C.byref(n))
    assert bad == 0, (bad, v.value, c.value, n.value)


def emu_chunk(states, chunks, i16, q16, step=1):
    """states: STREAM[n]; chunks: CHUNK[n]; i16 / q16: int16[n][len] -> (next states, out list, derot list, results)"""
    n = len(states)
    max_in = int(chunks['len_in'].max())
    nxt = np.zeros(n, STREAM)
    out_stride = int(max_in / float(chunks['resample'].min()) / 2) + 8
    out = np.zeros((n, out_stride), np.complex64)
    derot = np.zeros((n, max(max_in, 1)), np.complex64)
    res = np.zeros(n, RESULT)
    i16 = np.ascontiguousarray(i16, np.int16)
    q16 = np.ascontiguousarray(q16, np.int16)
    rc = emu().emu_fe_chunk(n, states.ctypes.data, nxt.ctypes.data, chunks.ctypes.data, i16.ctypes.data, q16.ctypes.data,
                            i16.shape[1], step, out.ctypes.data, out_stride, derot.ctypes.data, res.ctypes.data)
    assert rc == 0
    return nxt, [out[s, :res['len_out'][s]] for s in range(n)], [derot[s, :chunks['len_in'][s]] for s in range(n)], res


def fresh_states(n):
    st = np.zeros(n, STREAM)
    st['x1'] = -0.5
    return st


@pytest.mark.parametrize('resample', [0.5, 0.49999998, 0.50000003, 0.503, 0.61, 0.9])
def test_chunks_equal_the_oracle(resample):
    rng = np.random.default_rng(int(resample * 1e6) % 1000)
    n_streams, n_chunks = 3, 5
    ports = [O.PortFrontend() for _ in range(n_streams)]
    states = fresh_states(n_streams)
    theta = np.zeros((n_streams, 3), np.float64)
    for k in range(n_chunks):
        lens = rng.integers(1, 9000, n_streams) if k != 2 else np.array([2, 1030, 4096])
        chunks = np.zeros(n_streams, CHUNK)
        chunks['len_in'] = lens
        chunks['short_to_float'] = 1.0 / (1 << 14)
        chunks['c1'] = rng.uniform(-0.02, 0.02, n_streams)
        chunks['c2'] = rng.uniform(0.98, 1.02, n_streams)
        chunks['frequency_est_filtered'] = rng.choice([-1, 1], n_streams) * 10.0 ** rng.uniform(-8, -3, n_streams)
        chunks['phase_nco'] = rng.uniform(-3, 3, n_streams)
        chunks['resample'] = resample
        i16 = (rng.normal(0, 1500, (n_streams, lens.max())) + 40).astype(np.int16)
        q16 = (rng.normal(0, 1500, (n_streams, lens.max())) - 25).astype(np.int16)
        nxt, outs, derots, res = emu_chunk(states, chunks, i16, q16)
        for s in range(n_streams):
            c = chunks[s]
            po, pi, pd = ports[s].chunk(i16[s, :lens[s]], q16[s, :lens[s]], c['short_to_float'], c['c1'], c['c2'], c['frequency_est_filtered'],
                                        c['phase_nco'], float(c['resample']))
            rms = 1500.0 / (1 << 14)
            assert res['len_interp'][s] == len(pi) and res['len_out'][s] == len(po), (k, s, res[s], len(pi), len(po))
            assert np.abs(derots[s] - pd).max() <= 2e-6 * rms
            if len(po):
                assert np.abs(outs[s] - po).max() <= 2e-6 * rms, (k, s, np.abs(outs[s] - po).max())
            ps = ports[s].state[0]
            assert nxt['frequency_nco'][s].view(np.uint32) == ps['frequency_nco'].view(np.uint32)
            assert nxt['parity'][s] == ps['parity']
            assert abs(nxt['x1'][s] - ps['x1']) < 1e-5 and abs(nxt['dc_re'][s] - ps['dc_re']) < 1e-8
            assert np.abs(nxt['delay'][s] - ps['delay']).max() <= 2e-6 * rms and np.abs(nxt['hist'][s] - ps['hist']).max() <= 2e-6 * rms
            theta[s] += res['theta'][s]                            # the reference adds sample by sample into one float per call
            assert np.abs(theta[s] - ports[s].theta).max() <= 3e-5 * max(np.abs(ports[s].theta).max(), rms)     # theta1 is a cancelling sum, (k, s, theta[s], ports[s].theta)
        states = nxt


def test_chunk_with_more_nco_segments_than_the_plan_holds():
    """every step an exact tie (the decrement is half an ulp of the phase): no linear segments, the tiles plan for themselves
    and fall back to single steps"""
    rng = np.random.default_rng(9)
    states = fresh_states(2)
    states['frequency_nco'] = [1.0, -1.5]
    ports = [O.PortFrontend() for _ in range(2)]
    for s in range(2):
        ports[s].state['frequency_nco'] = states['frequency_nco'][s]
    chunks = np.zeros(2, CHUNK)
    chunks['len_in'], chunks['short_to_float'], chunks['c2'], chunks['resample'] = [3000, 2500], 2.0 ** -14, 1.0, 0.5
    chunks['frequency_est_filtered'] = [-2.0 ** -24, 2.0 ** -24]
    i16 = rng.normal(0, 1500, (2, 3000)).astype(np.int16)
    q16 = rng.normal(0, 1500, (2, 3000)).astype(np.int16)
    nxt, outs, derots, res = emu_chunk(states, chunks, i16, q16)
    for s in range(2):
        n = chunks['len_in'][s]
        po, _, pd = ports[s].chunk(i16[s, :n], q16[s, :n], 2.0 ** -14, 0.0, 1.0, float(chunks['frequency_est_filtered'][s]), 0.0, 0.5)
        assert np.abs(derots[s] - pd).max() <= 2e-6 * 0.09 and np.abs(outs[s] - po).max() <= 2e-6 * 0.09
        assert nxt['frequency_nco'][s].view(np.uint32) == ports[s].state[0]['frequency_nco'].view(np.uint32)


def test_cp_correlation_equals_the_oracle():
    rng = np.random.default_rng(3)
    for n, g in ((16384, 512), (32768, 256), (32768, 1024)):
        sym = (rng.normal(size=n + g) + 1j * rng.normal(size=n + g)).astype(np.complex64)
        sym[n:] = sym[:g] * np.exp(1j * 0.37) + 0.1 * (rng.normal(size=g) + 1j * rng.normal(size=g))
        sym = sym.astype(np.complex64)
        est = C.c_float()
        emu().emu_cp_correlate(sym.ctypes.data, n, g, C.byref(est))
        want = O.port_cp_correlate(sym, n, g)
        assert abs(est.value - want) <= 1e-5 * abs(want)
        assert abs(est.value * 2 * n - 0.37) < 0.03


def p1_test_signal(n, seed, p1_at=5000):
    """noise with a P1-like structure (C | A | B with the frequency-shifted repetitions of EN 302 755) at p1_at"""
    rng = np.random.default_rng(seed)
    x = (rng.normal(0, 0.02, n) + 1j * rng.normal(0, 0.02, n))
    a = (rng.normal(0, 0.1, 1024) + 1j * rng.normal(0, 0.1, 1024))
    sh = np.exp(2j * np.pi * np.arange(1024) / 1024)
    sym = np.concatenate([(a * sh)[:542], a, (a * sh)[542:]])
    x[p1_at:p1_at + 2048] += sym
    return x.astype(np.complex64)


def test_p1_correlator_equals_the_oracle_across_blocks():
    """the closed-form correlator against the oracle's delay lines and running sums, in one block and cut into blocks with the
    history handed over; 1e-5 of the peak (the reference's running sums drift in float)"""
    x = p1_test_signal(16000, 4)
    want, want_out = O.PortP1().correlate(x)
    peak = want.max()
    assert np.argmax(want) > 5000 and peak > 100 * np.median(want)
    for cuts in ([0, 16000], [0, 3000, 3001, 9500, 16000]):
        got = np.empty(len(x), np.float32)
        out = np.empty(len(x), np.complex64)
        for a, b in zip(cuts[:-1], cuts[1:]):
            hist = np.zeros(2046, np.complex64)
            h = x[max(0, a - 2046):a]
            if len(h):
                hist[-len(h):] = h
            blk = np.ascontiguousarray(x[a:b])
            c, o = np.empty(b - a, np.float32), np.empty(b - a, np.complex64)
            emu().emu_p1_correlate(blk.ctypes.data, b - a, hist.ctypes.data, a & 1023, c.ctypes.data, o.ctypes.data)
            got[a:b], out[a:b] = c, o
        assert np.abs(got - want).max() <= 1e-5 * peak
        assert np.abs(out - want_out).max() <= 1e-5 * np.sqrt(peak)
        assert np.argmax(got) == np.argmax(want)
