"""The front-end oracle (oracle/port/frontend_port.c) pinned to the reference: against tests/golden/frontend_ref.npz (what
the unmodified receiver computed for a window of chunks after lock, tools/make_golden_frontend.py) and, where the compiled
reference is present, against a live run of it with every sample of the window compared."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import fe_helpers as F


def run_port(w, keep, stream=F.STREAM):
    i16, q16, off = F.stream_input(w, stream)
    fe = O.PortFrontend()
    for k, v in w['state'].items():
        fe.state[k] = v
    for k, row in enumerate(w['info']):
        a, b = off[k], off[k + 1]
        out, interp, derot = fe.chunk(i16[a:b], q16[a:b], **F.chunk_args(row))
        assert len(interp) == int(row[F.COL['len_interp']])
        F.check_against(w, k, derot, out, keep)
        s = fe.state[0]
        # the carried loop state follows the reference's exactly (NCO) / to float rounding (DC average)
        assert np.float32(s['frequency_nco']) == np.float32(row[F.COL['frequency_nco_after']])
        assert abs(s['dc_re'] - row[F.COL['dc_re_after']]) < 1e-9 and abs(s['x1'] - row[F.COL['x1_after']]) < 1e-6
    return len(w['info'])


def test_port_equals_the_golden_window():
    assert run_port(F.load_golden(), F.KEEP) == F.COUNT - 1


@pytest.mark.skipif(not O.have_ref('libref_chain.so'), reason='compiled reference not present')
def test_port_equals_the_reference_live_and_the_fixture_is_current():
    t = F.run_reference()
    w = F.window(t)
    w['iq_sha'] = str(t['iq_sha'])
    assert run_port(w, 1) == F.COUNT - 1
    g = F.load_golden()
    assert g['iq_sha'] == w['iq_sha'] and np.array_equal(g['info'], w['info'])
    assert all(np.array_equal(a[::F.KEEP], b) for a, b in zip(w['decim'], g['decim']))


@pytest.mark.skipif(not O.have_ref('libref_chain.so'), reason='compiled reference not present')
def test_port_equals_the_reference_live_on_the_32k_stream():
    """the same on the 32K / 64 800 stream (tests/e2e_helpers.py 'c32e'): 40 chunks after lock, every sample"""
    t = F.run_reference('c32e', 330, 41)
    w = F.window(t, 330, 41)
    w['iq_sha'] = str(t['iq_sha'])
    assert (w['info'][:, F.COL['frequency_est_filtered']] != 0).any() and (w['info'][:, F.COL['phase_nco']] != 0).any()
    assert run_port(w, 1, 'c32e') == 40


def test_cp_correlation_port_recovers_a_known_rotation():
    rng = np.random.default_rng(0)
    n, g = 16384, 512
    sym = (rng.normal(size=n + g) + 1j * rng.normal(size=n + g)).astype(np.complex64)
    sym[n:] = sym[:g] * np.exp(-0.2j)
    est = O.port_cp_correlate(sym.astype(np.complex64), n, g)
    assert abs(est * 2 * n + 0.2) < 0.02


@pytest.mark.skipif(not O.have_ref('libref_chain.so'), reason='compiled reference not present')
def test_p1_correlator_port_equals_the_reference():
    """the oracle's restatement of p1_symbol's sliding correlator against the compiled reference's, sample by sample, on the
    head of the synthetic 16K stream (it holds a P1 symbol); 2e-6 of the peak (the reference is built -Ofast)"""
    from tests import e2e_helpers as H
    i16, q16, _, _ = H.make_stream('c16e')
    x = ((i16[:30000].astype(np.float32) + 1j * q16[:30000].astype(np.float32)) / 16384).astype(np.complex64)
    ref = O.ref_p1_trace(x)
    corr, _ = O.PortP1().correlate(x)
    assert ref.max() > 1000 * np.median(ref) and np.argmax(ref) == np.argmax(corr)
    assert np.abs(ref - corr).max() <= 2e-6 * ref.max()
