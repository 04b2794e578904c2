"""Front-end parity fixtures.  The unmodified reference receiver is run on the synthetic 16K stream of
tests/e2e_helpers.py ('c16e') with taps around its resampler and decimator (oracle/tap/DSP/*, oracle/ref_chain.cc): for a
window of chunks of dvbt2_demodulator::execute, taken after the receiver has locked (the frequency / phase / sample-rate
loops are running, one chunk has an odd resampler count), tests/golden/frontend_ref.npz holds the loop parameters of every
chunk, the carried state in front of the window and every 16th sample of what the reference computed (derotated samples and
decimator output).  The int16 input is regenerated from the modulator (its digest is checked)."""
import multiprocessing as mp
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'frontend_ref.npz')
STREAM, FIRST, COUNT, KEEP = 'c16e', 180, 30, 16
COL = {n: i for i, n in enumerate(['len_in', 'len_interp', 'resample', 'phase_nco', 'frequency_est_filtered', 'c1', 'c2',
                                   'frequency_nco_after', 'dc_re_after', 'dc_im_after', 'len_out', 'short_to_float', 'x1_after',
                                   'n_interp_after'])}
# samples: the reference is built -Ofast (its float rounding is not IEEE-ordered), the port and the GPU evaluate the DC
# average and the resampler phase in another order: agreement to 2e-6 of the signal RMS
TOL = 2e-6


def _ref_worker(path, stream=STREAM, first=FIRST, count=COUNT):
    from oracle import pyoracle as O
    from tests import e2e_helpers as H
    i16, q16, _, _ = H.make_stream(stream)
    rx = O.RefDemod(tap_fft=False, tap_frontend=(first, count))
    n = 3200000 if stream == STREAM else len(i16)      # enough for FIRST + COUNT chunks
    rx.feed(i16[:n], q16[:n])
    t = rx.frontend_taps()
    t['iq_sha'] = H.sha(i16) + H.sha(q16)
    np.savez(path, **t)


def run_reference(stream=STREAM, first=FIRST, count=COUNT):
    """-> the taps of oracle.pyoracle.RefDemod.frontend_taps for the window (fresh process: the reference keeps static state)"""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'taps.npz')
        p = mp.get_context('spawn').Process(target=_ref_worker, args=(path, stream, first, count))
        p.start()
        p.join()
        assert p.exitcode == 0
        g = np.load(path)
        return {k: g[k] for k in g.files}


def window(t, FIRST=FIRST, COUNT=COUNT):
    """Split the taps: chunk FIRST only provides the state in front of the window, chunks FIRST+1 .. are compared.
    -> dict(info [n][14], in_offset, state fields, derot / decim lists per chunk)"""
    info = t['info']
    assert len(info) >= FIRST + COUNT, len(info)
    off_in = np.concatenate([[0], np.cumsum(info[:, COL['len_in']]).astype(np.int64)])
    w = info[FIRST:FIRST + COUNT]
    od = np.concatenate([[0], np.cumsum(w[:, COL['len_in']]).astype(np.int64)])
    oi = np.concatenate([[0], np.cumsum(w[:, COL['len_interp']]).astype(np.int64)])
    oo = np.concatenate([[0], np.cumsum(w[:, COL['len_out']]).astype(np.int64)])
    d0, v0 = t['derot'][od[0]:od[1]], t['interp'][oi[0]:oi[1]]
    st = dict(dc_re=w[0, COL['dc_re_after']], dc_im=w[0, COL['dc_im_after']], frequency_nco=w[0, COL['frequency_nco_after']],
              x1=w[0, COL['x1_after']], delay=np.stack([d0[-1:-4:-1].real, d0[-1:-4:-1].imag], -1).astype(np.float32),
              hist=np.stack([v0[-63:].real, v0[-63:].imag], -1).astype(np.float32), parity=int(w[0, COL['n_interp_after']]) % 2)
    return dict(info=w[1:], in_offset=int(off_in[FIRST + 1]), state=st,
                derot=[t['derot'][od[k]:od[k + 1]] for k in range(1, COUNT)],
                decim=[t['decim'][oo[k]:oo[k + 1]] for k in range(1, COUNT)])


def save_golden(t):
    w = window(t)
    st = w['state']
    np.savez(GOLDEN, info=w['info'], in_offset=w['in_offset'], iq_sha=str(t['iq_sha']),
             st_scalars=np.array([st['dc_re'], st['dc_im'], st['frequency_nco'], st['x1'], st['parity']], np.float64),
             st_delay=st['delay'], st_hist=st['hist'],
             derot=np.concatenate([d[::KEEP] for d in w['derot']]), decim=np.concatenate([d[::KEEP] for d in w['decim']]))


def load_golden():
    """-> the window as window() returns it, with every KEEP-th sample of the reference's outputs"""
    g = np.load(GOLDEN)
    info = g['info']
    sc = g['st_scalars']
    st = dict(dc_re=sc[0], dc_im=sc[1], frequency_nco=sc[2], x1=sc[3], parity=int(sc[4]), delay=g['st_delay'], hist=g['st_hist'])
    nd = [len(range(0, int(n), KEEP)) for n in info[:, COL['len_in']]]
    no = [len(range(0, int(n), KEEP)) for n in info[:, COL['len_out']]]
    od, oo = np.concatenate([[0], np.cumsum(nd)]), np.concatenate([[0], np.cumsum(no)])
    return dict(info=info, in_offset=int(g['in_offset']), state=st, iq_sha=str(g['iq_sha']),
                derot=[g['derot'][od[k]:od[k + 1]] for k in range(len(info))],
                decim=[g['decim'][oo[k]:oo[k + 1]] for k in range(len(info))])


def stream_input(w, stream=STREAM):
    """the int16 I/Q of the window's chunks, regenerated by the test modulator -> (i16, q16, offsets per chunk)"""
    from tests import e2e_helpers as H
    i16, q16, _, _ = H.make_stream(stream)
    if 'iq_sha' in w:
        assert H.sha(i16) + H.sha(q16) == w['iq_sha'], 'the modulator no longer produces the stream the fixture was made from'
    off = w['in_offset'] + np.concatenate([[0], np.cumsum(w['info'][:, COL['len_in']]).astype(np.int64)])
    return i16, q16, off


def chunk_args(row):
    return dict(short_to_float=row[COL['short_to_float']], c1=row[COL['c1']], c2=row[COL['c2']],
                frequency_est_filtered=row[COL['frequency_est_filtered']], phase_nco=row[COL['phase_nco']],
                resample=row[COL['resample']])


def check_against(w, k, derot, out, keep):
    """chunk k of the window: computed derotated samples / decimator output against the reference's (every keep-th sample)"""
    rd, ro = w['derot'][k], w['decim'][k]
    rms = max(float(np.sqrt(np.mean(np.abs(rd) ** 2))), 1e-3)
    assert len(out[::keep]) == len(ro) and len(out) == int(w['info'][k, COL['len_out']]), (k, len(out), w['info'][k, COL['len_out']])
    if derot is not None:
        assert np.abs(derot[::keep] - rd).max() <= TOL * rms, (k, np.abs(derot[::keep] - rd).max() / rms)
    if len(ro):
        assert np.abs(out[::keep] - ro).max() <= TOL * rms, (k, np.abs(out[::keep] - ro).max() / rms)
