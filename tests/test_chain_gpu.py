"""Whole hot path on the GPU (FFT -> equalise -> TI -> demap -> LDPC -> BCH strip / descramble) through the C-ABI,
against the CPU oracle chain on the same synthetic IQ and against the transmitted BBFRAMEs.

Parity bars: BBFRAME bits, trial counts, int8 LLRs, TI cells, equalised cells and both feedback floats are compared
EXACTLY with the oracle chain fed the GPU's FFT output is not possible (the FFT has a tolerance), so the chain is
checked in two links: (a) GPU FFT within 1e-5 of the oracle DFT; (b) everything after the FFT bit-exact when both
sides start from the same frequency-domain symbols; (c) end to end from time samples: decoded bits identical to the
transmitted ones and to the oracle's."""
import numpy as np
import pytest

from oracle import pyoracle as O
from sdr_receiver_dvb_t2_b200 import engine as E
from sdr_receiver_dvb_t2_b200.chain import FrameChain
from tests.chain_helpers import port_receive
from tests.eq_helpers import tables
from tools.modulator import Modulator

pytestmark = pytest.mark.gpu


def test_config4_16k_64qam_r35_short_end_to_end(engine):
    """BASELINE config 4: 8 MHz 16K 64-QAM r3/5 short FECFRAME, full chain, TS-level parity"""
    import torch
    t = tables('c16')
    m = Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=96, ti_len=3, seed=4)
    frames = [m.frame(noise_cn_db=15.0) for _ in range(2)]
    ch = FrameChain(engine, t, mod=2, cod=1, fec_type=0, n_blocks=96, ti_len=3)
    time = torch.from_numpy(np.stack([f['time'] for f in frames])).cuda()
    r = ch.decode_frames(time, want_llr=True)
    engine.sync()
    bits = r['bits'].cpu().numpy()
    tl = r['trials_left'].cpu().numpy()
    assert (tl >= 0).all()
    sent = np.concatenate([f['bb'] for f in frames])
    assert np.array_equal(bits, sent)                                           # (c) what the transmitter sent
    # (c') and what the oracle chain decodes from the same IQ, group by group, with the same trial counts
    for fi, f in enumerate(frames):
        o = port_receive(t, m, f['time'])
        assert np.array_equal(bits[fi * 96:(fi + 1) * 96], o['bits'])
        for g, tr in enumerate(o['trials']):
            assert (tl[fi * 96 + 32 * g:fi * 96 + 32 * g + 32] == tr).all()
        # soft values agree except where the FFT's last-bit differences move a rounding: count them
        llr = r['llr'].cpu().numpy()[fi * 96:(fi + 1) * 96]
        d = np.abs(llr.astype(np.int16) - o['llr'].astype(np.int16))
        assert (d != 0).mean() < 2e-2 and d.max() <= 2
        assert np.allclose(r['sro'][fi], o['sro'], rtol=1e-3, atol=2e-2) and np.allclose(r['phase'][fi], o['phase'], atol=1e-3)


def test_post_fft_chain_bit_exact_c32_256qam(engine):
    """BASELINE config 1/2 geometry (32K ext PP7 GI 1/128, 256-QAM rotated r2/3, 202 FEC blocks, TI 67/67/68): with
    both sides starting from the SAME frequency-domain symbols every later value is bit-identical: cells, TI block,
    precision, LLRs, LDPC trial count and output bits of a lock-step group (which does not converge: the reference's
    wrapping LLR cast defeats 256-QAM on AWGN -- exactly reproduced)."""
    import torch
    t = tables('c32')
    p = t['p']
    m = Modulator(t, mod=3, cod=2, fec_normal=True, n_blocks=202, ti_len=3, seed=1)
    f = m.frame(noise_cn_db=19.5)
    ch = FrameChain(engine, t, mod=3, cod=2, fec_type=1, n_blocks=202, ti_len=3)
    time = torch.from_numpy(f['time'][None]).cuda()
    freq = engine.fft(time.reshape(p['len_frame'], p['fft_size']))               # GPU FFT output ...
    engine.sync()
    freq_h = freq.cpu().numpy()
    want = np.stack([O.port_fft(f['time'][i]) for i in (0, 1, 59)])
    assert np.abs(freq_h[[0, 1, 59]] - want).max() <= 1e-5 * np.abs(want).max()  # (a)
    # (b) ... is what BOTH chains continue from
    cells = []
    c, _, _ = O.port_equalize(0, freq_h[0], p['l_nulls'], p['k_total'], t['p2_map'], t['p2_ref'], t['h_odd_p2'], p['c_p2'], t['amp_p2'])
    cells.append(c)
    for i in range(1, 60):
        h = t['h_odd_data'] if i % 2 == 0 else t['h_even_data']
        c, _, _ = O.port_equalize(1, freq_h[i], p['l_nulls'], p['k_total'], t['data_map'][i - 1], t['data_ref'][i - 1], h,
                                  p['c_data'], t['amp_sp'], t['amp_cp'])
        cells.append(c)
    stream_o = np.concatenate(cells)[m.p2_start:m.p2_start + 202 * 8100]
    stream, sro, ph = ch.demodulate(time)
    engine.sync()
    assert np.array_equal(stream.cpu().numpy()[0].view(np.float32), stream_o.view(np.float32))
    perm = O.port_cell_permutation(68, 8100)
    ti_o = O.port_ti_blocks(stream_o, m.blocks, 8100, perm, [0, 0.0])
    ti = engine.ti_deinterleave(0, stream.reshape(-1), m.blocks)                 # before the demapper derotates it in place
    engine.sync()
    assert np.array_equal(ti.cpu().numpy().view(np.float32), ti_o.view(np.float32))
    r = ch.fec(stream, want_llr=True)
    engine.sync()
    off, llr_o, prec_o = 0, [], []
    for nf in m.blocks:
        l, s, pr, _ = O.port_demap(ti_o[off:off + nf * 8100], 3, 1, 1, 2)
        llr_o.append(l), prec_o.append(pr)
        off += nf * 8100
    llr_o = np.concatenate(llr_o)
    assert np.array_equal(r['precision'].cpu().numpy(), np.array(prec_o, np.float32))
    assert np.array_equal(r['llr'].cpu().numpy(), llr_o)
    # every lock-step group of the frame (6 x 32 + the trailing 10), not just the first
    tl, bits = r['trials_left'].cpu().numpy(), r['bits'].cpu().numpy()
    for g0 in range(0, 202, 32):
        g1 = min(g0 + 32, 202)
        tr, bits_o, _ = O.port_ldpc_decode(2, llr_o[g0:g1], 25)
        assert tr == -1 and (tl[g0:g1] == -1).all()                               # the reference drops every batch
        assert np.array_equal(bits[g0:g1], O.bch_strip_descramble(bits_o, 43200, 43040))


def test_c32_256qam_decodes_with_saturating_cast_option(engine):
    """the non-reference option T2B200_OPT_DEMAP_SATURATE makes the same frame decodable: all complete groups give
    back the transmitted BBFRAMEs"""
    import torch
    t = tables('c32')
    m = Modulator(t, mod=3, cod=2, fec_normal=True, n_blocks=202, ti_len=3, seed=2)
    f = m.frame(noise_cn_db=20.5)
    engine.set_option(E.OPT_DEMAP_SATURATE, 1)
    try:
        ch = FrameChain(engine, t, mod=3, cod=2, fec_type=1, n_blocks=202, ti_len=3)
        r = ch.decode_frames(torch.from_numpy(f['time'][None]).cuda())
        engine.sync()
    finally:
        engine.set_option(E.OPT_DEMAP_SATURATE, 0)
    tl = r['trials_left'].cpu().numpy()
    assert (tl >= 0).all()
    assert np.array_equal(r['bits'].cpu().numpy(), f['bb'])


def test_config5_multi_plp_r34_pooled_fec(engine):
    """BASELINE config 5 geometry on one GPU: 32K, two PLPs (type 1, contiguous), both 256-QAM rotated r3/4 64 800.
    The frame is demodulated once; each PLP runs its own time de-interleaver + demapper; the FEC blocks of BOTH PLPs are
    pooled into one lock-step LDPC call (what CodewordSharder spreads over the GPUs of a box) and every BBFRAME comes
    back as transmitted.  (Saturating LLR cast: the reference's wrapping cast cannot decode 256-QAM)"""
    import torch
    t = tables('c32')
    nb = (96, 64)
    m0 = Modulator(t, mod=3, cod=3, fec_normal=True, n_blocks=nb[0], ti_len=3, seed=21)
    m1 = Modulator(t, mod=3, cod=3, fec_normal=True, n_blocks=nb[1], ti_len=2, seed=22)
    bb1, _, _, stream1 = m1.plp_stream()
    f = m0.frame(noise_cn_db=22.5, extra_streams=[stream1])
    engine.set_option(E.OPT_DEMAP_SATURATE, 1)
    try:
        c0 = FrameChain(engine, t, mod=3, cod=3, fec_type=1, n_blocks=nb[0], ti_len=3, plp=0)
        c1 = FrameChain(engine, t, mod=3, cod=3, fec_type=1, n_blocks=nb[1], ti_len=2, plp=1, cell_offset=nb[0] * 8100)
        s0, _, _ = c0.demodulate(torch.from_numpy(f['time'][None]).cuda())
        s1 = c1.plp_cells(c0.last_cells)
        llrs = []
        for ch, s in ((c0, s0), (c1, s1)):
            ti = engine.ti_deinterleave(ch.plp, s.reshape(-1), ch.blocks)
            llrs.append(engine.demap(ti, ch.blocks, 3, 1, 1, 3)['llr'])
        pooled = torch.cat(llrs)                                  # 160 codewords = 5 lock-step groups
        r = engine.ldpc_decode(c0.code, pooled, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE)
        engine.sync()
    finally:
        engine.set_option(E.OPT_DEMAP_SATURATE, 0)
    bits = r['bits'].cpu().numpy()
    assert (r['trials_left'].cpu().numpy() >= 0).all()
    assert np.array_equal(bits[:nb[0]], f['bb']) and np.array_equal(bits[nb[0]:], bb1)


def test_fused_frames_call_equals_the_staged_chain(engine):
    """t2b200_frames_decode (one C call, everything chained on the device) gives exactly what the per-stage calls give:
    bits, trial counts, both feedback floats of every symbol, SNR -- config 4 geometry, device and host buffers"""
    import torch
    t = tables('c16')
    m = Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=96, ti_len=3, seed=9)
    frames = [m.frame(noise_cn_db=15.5) for _ in range(3)]
    time = np.stack([f['time'] for f in frames])
    ch = FrameChain(engine, t, mod=2, cod=1, fec_type=0, n_blocks=96, ti_len=3)
    a = ch.decode_frames(torch.from_numpy(time).cuda())
    b = ch.decode_frames_fused(torch.from_numpy(time).cuda())
    engine.sync()
    assert np.array_equal(a['bits'].cpu().numpy(), b['bits'].cpu().numpy())
    assert np.array_equal(a['trials_left'].cpu().numpy(), b['trials_left'].cpu().numpy())
    assert np.array_equal(a['sro'], b['sro'].cpu().numpy()) and np.array_equal(a['phase'], b['phase'].cpu().numpy())
    assert np.array_equal(a['snr'], b['snr'].cpu().numpy())
    assert np.array_equal(b['bits'].cpu().numpy(), np.concatenate([f['bb'] for f in frames]))
    c = ch.decode_frames_fused(time)                             # pageable host buffers in and out
    assert np.array_equal(c['bits'], b['bits'].cpu().numpy()) and np.array_equal(c['sro'], a['sro'])


def test_frame_closing_symbol_geometry_fused_and_staged(engine):
    """a mode WITH a frame-closing symbol (32K, PP that needs one): the PLP runs into the FC symbol's cells; the one-call
    pipeline and the staged chain agree and return the transmitted BBFRAMEs"""
    import torch
    t = tables('c32fc')
    p = t['p']
    assert p['l_fc'] == 1
    # enough FEC blocks that the PLP reaches the frame-closing symbol
    nd = p['len_frame'] - p['n_p2'] - p['l_fc']
    cap = p['c_p2'] - 2200 + nd * p['c_data'] + p['n_fc']
    nb = min(cap // 10800, 8 * (cap // 10800 // 8) + 5)
    m = Modulator(t, mod=2, cod=1, fec_normal=True, n_blocks=nb, ti_len=2, seed=31)
    assert nb * 10800 > p['c_p2'] - 2200 + nd * p['c_data']          # ... it does
    f = m.frame(noise_cn_db=17.0)
    ch = FrameChain(engine, t, mod=2, cod=1, fec_type=1, n_blocks=nb, ti_len=2)
    x = torch.from_numpy(f['time'][None]).cuda()
    a = ch.decode_frames(x)
    b = ch.decode_frames_fused(x)
    engine.sync()
    full = (nb // 32) * 32
    assert np.array_equal(a['bits'].cpu().numpy(), b['bits'].cpu().numpy())
    assert np.array_equal(a['sro'], b['sro'].cpu().numpy()) and np.array_equal(a['phase'], b['phase'].cpu().numpy())
    tl = b['trials_left'].cpu().numpy()
    assert (tl >= 0).all()
    assert np.array_equal(b['bits'].cpu().numpy(), f['bb']) and full >= 0
