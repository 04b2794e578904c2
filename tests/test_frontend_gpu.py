"""N2 on the GPU: t2b200_frontend_execute / t2b200_cp_correlate through the C-ABI against the oracle port of the
reference's front-end (oracle/port/frontend_port.c, itself pinned to the compiled reference by tests/test_oracle_frontend.py)
and against tests/golden/frontend_ref.npz (what the unmodified receiver computed after lock).

Bars: chunk lengths, resampler counts, decimator phase and the NCO phase carried between chunks are exact; samples agree to
2e-6 of the signal RMS, the IQ-imbalance statistics to 3e-5 (float sums in another order than the reference's serial ones)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from sdr_receiver_dvb_t2_b200 import engine as E
from tests import fe_helpers as F

pytestmark = pytest.mark.gpu


def run_gpu_vs_port(engine, n_streams, n_chunks, resample, seed, device_buffers=False, step=1):
    rng = np.random.default_rng(seed)
    engine.frontend_configure(n_streams, 40000)
    ports = [O.PortFrontend() for _ in range(n_streams)]
    theta = np.zeros((n_streams, 3))
    rms = 1500.0 / (1 << 14)
    for k in range(n_chunks):
        lens = rng.integers(1, 36000, n_streams) if k != 1 else np.resize(np.array([2, 1030, 4096, 0, 1]), n_streams)
        chunks = np.zeros(n_streams, E.FE_CHUNK)
        chunks['len_in'] = lens
        chunks['short_to_float'] = 1.0 / (1 << 14)
        chunks['c1'] = rng.uniform(-0.02, 0.02, n_streams)
        chunks['c2'] = rng.uniform(0.98, 1.02, n_streams)
        chunks['frequency_est_filtered'] = rng.choice([-1, 1], n_streams) * 10.0 ** rng.uniform(-8, -3, n_streams)
        chunks['phase_nco'] = rng.uniform(-3, 3, n_streams)
        chunks['resample'] = resample
        width = int(lens.max()) * step + 3
        iq = (rng.normal(0, 1500, (2, n_streams, width)) + 30).astype(np.int16)
        if step == 2:                                           # airspy layout: I and Q interleaved (convert_input = 2)
            i_in, q_in = iq[0], iq[0].reshape(-1)[1:]
            q_view = np.lib.stride_tricks.as_strided(iq[0].reshape(-1)[1:], (n_streams, width - 1), (width * 2, 2))
        else:
            i_in, q_in, q_view = iq[0], iq[1], iq[1]
        if device_buffers:
            import torch
            ti = torch.from_numpy(iq).cuda()
            gi, gq = ti[0], (ti[0].reshape(-1)[1:] if step == 2 else ti[1])
            out, res = engine.frontend_execute(gi, gq, chunks, stream_stride=width, sample_step=step)
            out = out.cpu().numpy()
        else:
            out, res = engine.frontend_execute(i_in, q_in, chunks, stream_stride=width, sample_step=step)
        for s in range(n_streams):
            c = chunks[s]
            po, pi, _ = ports[s].chunk(i_in[s, :lens[s] * step], q_view[s, :lens[s] * step], c['short_to_float'], c['c1'], c['c2'],
                                       c['frequency_est_filtered'], c['phase_nco'], float(c['resample']), stride=step)
            assert res['len_interp'][s] == len(pi) and res['len_out'][s] == len(po), (k, s, res[s], len(pi), len(po))
            if len(po):
                assert np.abs(out[s, :len(po)] - po).max() <= F.TOL * rms, (k, s, np.abs(out[s, :len(po)] - po).max() / rms)
            st, ps = engine.frontend_state(s), ports[s].state[0]
            assert np.float32(st['frequency_nco']).view(np.uint32) == np.float32(ps['frequency_nco']).view(np.uint32)
            assert st['parity'] == ps['parity'] and abs(st['x1'] - ps['x1']) < 1e-5 and abs(st['dc_re'] - ps['dc_re']) < 1e-8
            assert np.abs(st['delay'] - ps['delay']).max() <= F.TOL * rms and np.abs(st['hist'] - ps['hist']).max() <= F.TOL * rms
            theta[s] += res['theta'][s]
            assert np.abs(theta[s] - ports[s].theta).max() <= 3e-5 * max(np.abs(ports[s].theta).max(), rms)     # theta1 is a cancelling sum


@pytest.mark.parametrize('resample', [0.5, 0.49999998, 0.50000003, 0.503, 0.61, 0.9])
def test_chunks_equal_the_oracle(engine, resample):
    run_gpu_vs_port(engine, 5, 4, resample, seed=int(resample * 1e6) % 997)


def test_device_buffers_and_interleaved_input(engine):
    run_gpu_vs_port(engine, 3, 3, 0.5, seed=5, device_buffers=True)
    run_gpu_vs_port(engine, 3, 3, 0.5, seed=6, step=2)
    run_gpu_vs_port(engine, 2, 2, 0.52, seed=7, device_buffers=True, step=2)


def test_golden_window_of_the_reference_receiver(engine):
    """stream 0 replays the window the reference receiver went through after lock (its loop values, its carried state);
    stream 1 runs random data next to it"""
    w = F.load_golden()
    i16, q16, off = F.stream_input(w)
    engine.frontend_configure(2, 20000)
    st = np.zeros(1, E.FE_STATE)[0]
    for k, v in w['state'].items():
        st[k] = v
    engine.frontend_set_state(0, st)
    rng = np.random.default_rng(1)
    for k, row in enumerate(w['info']):
        a, b = int(off[k]), int(off[k + 1])
        n = b - a
        iq = np.zeros((2, 2, n), np.int16)
        iq[0, 0], iq[1, 0] = i16[a:b], q16[a:b]
        iq[:, 1] = rng.normal(0, 900, (2, n)).astype(np.int16)
        chunks = np.zeros(2, E.FE_CHUNK)
        for f, v in F.chunk_args(row).items():
            chunks[f] = v
        chunks['len_in'] = n
        out, res = engine.frontend_execute(iq[0], iq[1], chunks)
        assert res['len_interp'][0] == int(row[F.COL['len_interp']])
        F.check_against(w, k, None, out[0, :res['len_out'][0]], F.KEEP)
        s = engine.frontend_state(0)
        assert np.float32(s['frequency_nco']) == np.float32(row[F.COL['frequency_nco_after']])
        assert abs(s['dc_re'] - row[F.COL['dc_re_after']]) < 1e-8 and abs(s["x1"] - row[F.COL["x1_after"]]) < 1e-6


def test_chunk_with_more_nco_segments_than_the_plan_holds(engine):
    """every step an exact tie (the decrement is half an ulp of the phase): no linear segments, the tiles plan for themselves
    and fall back to single steps"""
    rng = np.random.default_rng(9)
    engine.frontend_configure(2, 4000)
    ports = [O.PortFrontend() for _ in range(2)]
    for s, v in enumerate((1.0, -1.5)):
        st = engine.frontend_state(s)
        st['frequency_nco'] = v
        engine.frontend_set_state(s, st)
        ports[s].state['frequency_nco'] = v
    chunks = np.zeros(2, E.FE_CHUNK)
    chunks['len_in'], chunks['short_to_float'], chunks['c2'], chunks['resample'] = [3000, 2500], 2.0 ** -14, 1.0, 0.5
    chunks['frequency_est_filtered'] = [-2.0 ** -24, 2.0 ** -24]
    iq = rng.normal(0, 1500, (2, 2, 3000)).astype(np.int16)
    out, res = engine.frontend_execute(iq[0], iq[1], chunks)
    for s in range(2):
        n = chunks['len_in'][s]
        po, _, _ = ports[s].chunk(iq[0, s, :n], iq[1, s, :n], 2.0 ** -14, 0.0, 1.0, float(chunks['frequency_est_filtered'][s]), 0.0, 0.5)
        assert res['len_out'][s] == len(po) and np.abs(out[s, :len(po)] - po).max() <= F.TOL * 0.09
        assert np.float32(engine.frontend_state(s)['frequency_nco']).view(np.uint32) == ports[s].state[0]['frequency_nco'].view(np.uint32)


def test_cp_correlation_equals_the_oracle(engine):
    rng = np.random.default_rng(3)
    for n, g in ((16384, 512), (32768, 256), (32768, 1024)):
        sym = (rng.normal(size=(7, n + g)) + 1j * rng.normal(size=(7, n + g))).astype(np.complex64)
        sym[:, n:] = sym[:, :g] * np.exp(1j * rng.uniform(-1, 1, (7, 1))) + 0.1 * rng.normal(size=(7, g))
        sym = np.ascontiguousarray(sym.astype(np.complex64))
        est = engine.cp_correlate(sym, n, g)
        want = np.array([O.port_cp_correlate(sym[i], n, g) for i in range(7)])
        assert np.abs(est - want).max() <= 1e-5 * np.abs(want).max()


def test_p1_correlator_equals_the_oracle_across_blocks(engine):
    """t2b200_p1_correlate against the oracle's delay lines and running sums (p1_symbol.cpp:75-178), one block and several
    blocks with the history handed over, host and device buffers; 1e-5 of the peak"""
    import torch
    from tests.test_frontend_emu import p1_test_signal
    x = p1_test_signal(40000, 4, p1_at=21000)
    want, want_out = O.PortP1().correlate(x)
    peak = want.max()
    for cuts, dev in (([0, 40000], False), ([0, 3000, 3001, 22000, 40000], True)):
        got = np.empty(len(x), np.float32)
        out = np.empty(len(x), np.complex64)
        for a, b in zip(cuts[:-1], cuts[1:]):
            hist = np.zeros(2046, np.complex64)
            h = x[max(0, a - 2046):a]
            if len(h):
                hist[-len(h):] = h
            blk = np.ascontiguousarray(x[a:b])
            if dev:
                c, o = engine.p1_correlate(torch.from_numpy(blk).cuda(), torch.from_numpy(hist).cuda(), a & 1023)
                engine.sync()                                     # device buffers keep the call asynchronous (include/t2b200.h)
                c, o = c.cpu().numpy(), o.cpu().numpy()
            else:
                c, o = engine.p1_correlate(blk, hist if a else None, a & 1023)
            got[a:b], out[a:b] = c, o
        assert np.abs(got - want).max() <= 1e-5 * peak and np.abs(out - want_out).max() <= 1e-5 * np.sqrt(peak)
        assert np.argmax(got) == np.argmax(want)


def test_reset_and_bad_arguments(engine):
    engine.frontend_configure(2, 5000)
    chunks = np.zeros(2, E.FE_CHUNK)
    chunks['len_in'], chunks['short_to_float'], chunks['c2'], chunks['resample'] = 4000, 2.0 ** -14, 1.0, 0.5
    iq = np.random.default_rng(0).normal(0, 1000, (2, 2, 4000)).astype(np.int16)
    out1, res1 = engine.frontend_execute(iq[0], iq[1], chunks)
    out2, _ = engine.frontend_execute(iq[0], iq[1], chunks)
    assert not np.array_equal(out1[:, :64], out2[:, :64])           # the second chunk starts from the first one's tail
    engine.frontend_reset()
    out3, res3 = engine.frontend_execute(iq[0], iq[1], chunks)
    assert np.array_equal(out1, out3) and np.array_equal(res1, res3)
    engine.frontend_reset(1)                                           # one stream only
    out4, _ = engine.frontend_execute(iq[0], iq[1], chunks)
    assert np.array_equal(out4[1], out1[1]) and np.array_equal(out4[0], out2[0])
    bad = chunks.copy()
    bad['len_in'] = 6000
    with pytest.raises(E.T2Error):
        engine.frontend_execute(iq[0], iq[1], bad)
    bad = chunks.copy()
    bad['resample'] = 1.5
    with pytest.raises(E.T2Error):
        engine.frontend_execute(iq[0], iq[1], bad)
    with pytest.raises(E.T2Error):
        engine.frontend_execute(iq[0], iq[1], chunks, out=np.zeros((2, 100), np.complex64))
    with pytest.raises(E.T2Error):
        engine.frontend_configure(0, 100)
    st = engine.frontend_state(0)
    for field, v in (('x1', -3.0), ('parity', 2), ('frequency_nco', 7.0)):
        bad_st = st.copy()
        bad_st[field] = v
        with pytest.raises(E.T2Error):
            engine.frontend_set_state(0, bad_st)
