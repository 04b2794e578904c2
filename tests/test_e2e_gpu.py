"""End-to-end parity of the GPU chain against the reference's OWN receiver (north_star: "output TS bytes
match the reference bit-exactly on the same synthetic IQ").

The unmodified reference (oracle/_ref/libref_chain.so, CPU) receives the synthetic int16 I/Q stream -- P1 detection, L1
parsing, resampler, symbol framing, FFTW, equaliser, TI, demapper, LDPC, BCH stage, bb_de_header -- on THIS box; its TS must
equal the committed golden digest.  The GPU chain (t2b200_frames_decode + t2b200_ts_packetize) then starts from the FFT
windows the reference cut out of the stream (replay mode) and must return the same BBFRAMEs and the same TS
datagrams byte for byte.  Configs: BASELINE config 4 (16K 64-QAM r3/5 short) and 32K 64-QAM r3/5 with 64 800-bit frames."""
import numpy as np
import pytest

from oracle import pyoracle as O
from sdr_receiver_dvb_t2_b200 import engine as E
from sdr_receiver_dvb_t2_b200.chain import FrameChain
from tests import e2e_helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', sorted(H.CONFIGS))
def test_gpu_chain_ts_equals_the_reference_receiver(engine, name):
    import torch
    assert O.have_ref('libref_chain.so'), 'oracle/_ref/libref_chain.so must travel to the GPU box (built by __graft_entry__.build)'
    c, g = H.CONFIGS[name], H.golden()[name]
    t = H.run_reference(name)                                  # the reference itself, on this box's CPU
    assert t['iq_sha'] == g['iq_sha'] and H.sha(t['ts']) == g['ts_sha'] and len(t['bb_bits']) == g['n_bbframes']
    frames = H.decoded_frames(t)
    tx = H.transmitter(name)
    ch = FrameChain(engine, tx.tables, mod=c['mod'], cod=c['cod'], fec_type=int(c['fec_normal']), n_blocks=c['n_blocks'],
                    ti_len=c['ti_len'], l1_post_size=tx.l1_post_size)
    d_frames = torch.from_numpy(frames).cuda()
    for fused in (True, False):
        r = ch.decode_frames_fused(d_frames) if fused else ch.decode_frames(d_frames, host_feedback=False)
        engine.sync()
        assert (r['trials_left'].cpu().numpy() >= 0).all()
        bits = r['bits'].cpu().numpy()
        assert np.array_equal(bits, t['bb_bits']), 'BBFRAMEs differ from the reference bch_decoder output'
        engine.ts_reset(0)
        ts, dl, st = engine.ts_packetize(r['bits'])
        engine.sync()
        assert (st == 0).all()
        assert np.array_equal(dl, t['ts_datagrams'])
        ts = ts.cpu().numpy()
        assert np.array_equal(ts, t['ts']), 'TS differs from the reference bb_de_header datagrams'
        assert H.sha(ts) == g['ts_sha']
    # SNR estimate of the demapper (llr_demapper.cpp:659-660), one per TI block, against the values the reference emitted
    snr = r['snr'].cpu().numpy() if hasattr(r['snr'], 'cpu') else np.asarray(r['snr'])
    assert np.allclose(snr[:len(t['snr'])], t['snr'][:len(snr)], atol=0.02)


def test_int16_windows_and_packed_bits_give_the_golden_ts(engine):
    """The front-end-facing form of the frame pipeline -- int16 I/Q in (converted while the FFT loads them: the first step of
    dvbt2_demodulator::execute, dvbt2_demodulator.cpp:182-186), packed BBFRAME bits out (t2b200_frames_decode_i16 +
    T2B200_LDPC_PACK_BITS) -- with the FFT windows cut straight out of the int16 stream at the transmitter's own symbol
    positions (ideal synchronisation).  No reference at run time: the TS must hash to the committed golden of the
    reference's receiver, i.e. the same BBFRAMEs come out although the windows differ from the ones the reference's
    resampler / P1 timing produced."""
    import torch
    name = 'c32e'
    c, g = H.CONFIGS[name], H.golden()[name]
    i16, q16, bb, tx = H.make_stream(name)
    assert H.sha(i16) + H.sha(q16) == g['iq_sha']
    m = tx.m
    N, gi, L = m.fft_size, m.guard_interval_size, m.len_frame
    frame_len = 2048 + L * (N + gi)
    first_frame = g['first_bbframe'] // c['n_blocks']
    n_frames = g['n_bbframes'] // c['n_blocks']
    win = np.empty((n_frames, L, N, 2), np.int16)
    for f in range(n_frames):
        s0 = 4096 + (first_frame + f) * frame_len + 2048                # lead-in + frames before + P1
        for l in range(L):
            a = s0 + l * (N + gi) + gi
            win[f, l, :, 0], win[f, l, :, 1] = i16[a:a + N], q16[a:a + N]
    ch = FrameChain(engine, tx.tables, mod=c['mod'], cod=c['cod'], fec_type=int(c['fec_normal']), n_blocks=c['n_blocks'],
                    ti_len=c['ti_len'], l1_post_size=tx.l1_post_size)
    flags = E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE | E.LDPC_PACK_BITS
    for dev in (False, True):                                           # host int16 buffer, then a device one
        x = torch.from_numpy(win).cuda() if dev else win
        r = ch.decode_frames_fused(x, flags=flags, scale=1.0 / (1 << 14))
        engine.sync()
        packed = r['bits'].cpu().numpy() if dev else r['bits']
        tl = r['trials_left'].cpu().numpy() if dev else r['trials_left']
        assert (tl >= 0).all()
        bits = np.unpackbits(packed, axis=1)
        assert np.array_equal(bits, bb[g['first_bbframe']:g['first_bbframe'] + g['n_bbframes']])
        engine.ts_reset(0)
        ts, dl, st = engine.ts_packetize(np.ascontiguousarray(bits))
        assert H.sha(ts) == g['ts_sha'] and len(ts) == g['ts_bytes']
