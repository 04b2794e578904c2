"""a2: the native frequency de-interleaver address tables (t2b200_freq_deinterleaver_table, host-side, EN 302 755
8.5) equal the tables the compiled reference's address_freq_deinterleaver built (fixtures tests/golden/tables_*.npz, dumped
by tools/make_golden_tables.py) for P2, data and frame-closing symbols in 16K and 32K."""
import numpy as np
import pytest

from sdr_receiver_dvb_t2_b200 import engine as E
from tests.eq_helpers import tables


@pytest.mark.parametrize('fixture', ['c16', 'c32', 'c32fc'])
def test_native_tables_equal_the_reference_fixtures(fixture):
    t = tables(fixture)
    p = t['p']
    kinds = [('p2', p['c_p2']), ('data', p['c_data'])] + ([('fc', p['n_fc'])] if p['l_fc'] else [])
    for k, n in kinds:
        e, o = E.freq_deinterleaver_table(p['fft_size'], n)
        assert np.array_equal(e, t['h_even_' + k][:n]) and np.array_equal(o, t['h_odd_' + k][:n]), (fixture, k)
        assert sorted(e) == list(range(n)) and sorted(o) == list(range(n))          # permutations


def test_rejects_what_it_cannot_build():
    with pytest.raises(E.T2Error):
        E.freq_deinterleaver_table(8192, 6000)          # only the one-P2-symbol FFT sizes (16K, 32K) are built
    with pytest.raises(E.T2Error):
        E.freq_deinterleaver_table(32768, 0)
