"""Pin the CPU oracle of the FFT convention and the equaliser (oracle/port/eq_port.c) against the compiled
reference.  The reference is built -Ofast (reciprocal maths, its own sqrt), so cells agree to a relative
tolerance (<= 1e-3; observed ~5e-7, a sin/cos table step where an index rounds differently);
the two feedback floats and the table fixtures must agree exactly / to float rounding."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests.eq_helpers import kind_tables, synth_symbol, tables

needs_ref = pytest.mark.skipif(not O.have_ref('libref_chain.so'), reason='reference chain not built here')


def test_atan2_approx_properties():
    L = O.port()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        y, x = rng.standard_normal(2).astype(np.float32)
        assert abs(L.port_atan2_approx(y, x) - np.arctan2(y, x)) < 2e-3          # the polynomial's own error
    assert L.port_atan2_approx(1.0, 0.0) == np.float32(np.pi / 2) and L.port_atan2_approx(0.0, -1.0) == np.float32(-np.pi)


def test_port_fft_is_shifted_unnormalised_dft():
    rng = np.random.default_rng(2)
    for n in (1024, 16384, 32768):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        want = np.fft.fftshift(np.fft.fft(x.astype(np.complex128)))
        got = O.port_fft(x)
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


@needs_ref
def test_fixture_tables_and_equaliser_match_reference_c32():
    """one receiver instance per process: everything that needs the live reference for C32 lives here"""
    t = tables('c32')
    rx = O.RefRx('32K', True, 7, '1/128', 59)
    live = rx.tables()
    assert rx.p == t['p']
    for k in ('data_map', 'data_ref', 'p2_map', 'p2_ref'):
        assert np.array_equal(live[k], t[k]), k
    for k in ('h_even_data', 'h_odd_data', 'h_even_p2', 'h_odd_p2'):
        n = len(t[k])
        assert np.array_equal(live[k][:n], t[k]), k
    p = t['p']
    rng = np.random.default_rng(3)
    # FFT wrapper: FFTW single precision vs the port's double-precision DFT
    x = (rng.standard_normal(p['fft_size']) + 1j * rng.standard_normal(p['fft_size'])).astype(np.complex64)
    a, b = rx.fft(x), O.port_fft(x)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    for idx in (1, 2, 3, 4, 58):
        f = synth_symbol(t, t['data_map'][idx - 1], t['data_ref'][idx - 1], rng)
        ref_cells, ref_sro, ref_ph = rx.data_symbol(idx, f)
        h = t['h_odd_data'] if idx % 2 == 0 else t['h_even_data']
        cells, sro, ph = O.port_equalize(1, f, p['l_nulls'], p['k_total'], t['data_map'][idx - 1], t['data_ref'][idx - 1], h,
                                         p['c_data'], t['amp_sp'], t['amp_cp'])
        err = np.abs(cells - ref_cells) / np.maximum(np.abs(ref_cells), 1e-2)
        assert err.max() < 1e-3 and np.median(err) < 1e-6
        assert abs(sro - ref_sro) <= 1e-4 * max(1.0, abs(ref_sro)) and abs(ph - ref_ph) < 1e-5
    f = synth_symbol(t, t['p2_map'], t['p2_ref'], rng)
    ref_cells, ref_sro, ref_ph, _ = rx.p2_symbol(f)
    cells, sro, ph = O.port_equalize(0, f, p['l_nulls'], p['k_total'], t['p2_map'], t['p2_ref'], t['h_odd_p2'], p['c_p2'], t['amp_p2'])
    err = np.abs(cells - ref_cells) / np.maximum(np.abs(ref_cells), 1e-2)
    assert err.max() < 1e-3 and np.median(err) < 1e-6
    assert abs(sro - ref_sro) <= 1e-4 * max(1.0, abs(ref_sro)) and abs(ph - ref_ph) < 1e-5
