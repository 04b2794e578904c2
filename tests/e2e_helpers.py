"""End-to-end parity fixtures: synthetic 8 MHz DVB-T2 signals (P1 + L1 signalling + guard intervals, int16 I/Q at 64/7 MHz,
tools/modulator.py::Transmitter) pushed through the UNMODIFIED reference receiver -- dvbt2_demodulator::execute ->
symbol_acquisition -> p2/data symbol -> time_deinterleaver -> llr_demapper -> ldpc_decoder -> bch_decoder -> bb_de_header
(dvbt2_demodulator.cpp:145-448 and the signal chain behind it), compiled into oracle/_ref/libref_chain.so.  The TS it emits
is the golden (tests/golden/e2e_ref.json holds its digest, made by tools/make_golden_e2e.py); the FFT windows it cut out of
the sample stream (`in_fft`, dvbt2_demodulator.cpp:332) are what the replay-mode chains under test start from.

Modes: the guard interval 1/32 + PP4 variants lock within three frames, and every TI block holds a
multiple of 32 FEC blocks: llr_demapper keeps the PLP ids of a 32-frame batch in a stack array that only survives inside one
call (llr_demapper.cpp:369,510-512), so a batch that straddles two TI blocks makes the reference read garbage."""
import hashlib
import json
import multiprocessing as mp
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {
    # BASELINE config 4: 8 MHz 16K 64-QAM r3/5, 16 200-bit FECFRAMEs
    'c16e': dict(fft='16K', pp=4, gi='1/32', n_data=20, mod=2, cod=1, fec_normal=False, n_blocks=96, ti_len=3, cn_db=16.0,
                 seed=4, n_frames=7),
    # 32K, 64-QAM r3/5, 64 800-bit FECFRAMEs: the 32K / normal-frame mode the reference decodes on AWGN
    'c32e': dict(fft='32K', pp=4, gi='1/32', n_data=26, mod=2, cod=1, fec_normal=True, n_blocks=64, ti_len=2, cn_db=19.0,
                 seed=5, n_frames=7),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def transmitter(name):
    from sdr_receiver_dvb_t2_b200 import engine as E
    from tools.modulator import Transmitter
    c = CONFIGS[name]
    m = E.mode_init(c['fft'], True, c['pp'], c['gi'], c['n_data'])
    return Transmitter(m, mod=c['mod'], cod=c['cod'], fec_normal=c['fec_normal'], n_blocks=c['n_blocks'], ti_len=c['ti_len'],
                       seed=c['seed'])


def make_stream(name):
    """-> (I int16, Q int16, transmitted BBFRAME bits [n_frames * n_blocks][K_bch], Transmitter)"""
    tx = transmitter(name)
    c = CONFIGS[name]
    i16, q16, frames = tx.stream(c['n_frames'], cn_db=c['cn_db'])
    return i16, q16, np.concatenate([f['bb'] for f in frames]), tx


def _ref_worker(name, path):
    from oracle import pyoracle as O
    i16, q16, _, _ = make_stream(name)
    rx = O.RefDemod()
    rx.feed(i16, q16)
    t = rx.taps()
    info = t['fft_info']
    # whole frames received with the FEC chain running: P2 (kind 1) with the deint_start flag, then its data symbols
    t['params'] = json.dumps(rx.params())
    t['iq_sha'] = sha(i16) + sha(q16)
    np.savez(path, **t)


def run_reference(name):
    """The reference's own receiver over the stream of `name` (fresh process: its stages keep static state).
    -> dict: ts, ts_datagrams, bb_bits [n][K_bch], fft_in [n_symbols][fft_size], fft_info [n_symbols][3], params, iq_sha"""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'taps.npz')
        p = mp.get_context('spawn').Process(target=_ref_worker, args=(name, path))
        p.start()
        p.join()
        assert p.exitcode == 0, 'reference receiver died (exit code %s)' % p.exitcode
        g = np.load(path)
        t = {k: g[k] for k in g.files}
    t['params'] = json.loads(str(t['params']))
    t['iq_sha'] = str(t['iq_sha'])
    n = len(t['bb_len'])
    t['bb_bits'] = t['bb_bits'].reshape(n, -1) if n else t['bb_bits']
    return t


def decoded_frames(t):
    """FFT windows of the T2 frames the reference ran its FEC chain on: complex64 [F][len_frame][fft_size]"""
    info, L = t['fft_info'], t['params']['len_frame']
    starts = [i for i in range(len(info)) if info[i, 0] == 1 and (info[i, 2] & 1) and i + L <= len(info)]
    # (the P2 symbol on which the chain STARTS already has deint_start clear when its FFT runs, so it is not in the list,
    # yet its cells do go into the de-interleaver: dvbt2_demodulator.cpp:376-388)
    first = [i for i in range(len(info)) if info[i, 0] == 1 and info[i, 2] == 2 and i + L <= len(info)][-1:]
    starts = first + starts
    for s in starts:
        assert (info[s + 1:s + L, 0] == 2).all() and (info[s + 1:s + L, 1] == np.arange(1, L)).all()
    return np.stack([t['fft_in'][s:s + L] for s in starts])


def golden():
    return json.load(open(os.path.join(ROOT, 'tests', 'golden', 'e2e_ref.json')))
