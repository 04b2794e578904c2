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
