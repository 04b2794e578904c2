"""a2: the native mode / pilot-table builder (t2b200_mode_init, t2b200_pilot_tables: csrc/pilots.cpp) against the
tables the unmodified reference builds (dvbt2_definition.cpp:20-648, pilot_generator.cpp:69-2166) for every SISO
16K / 32K mode: digests made by tools/make_golden_pilots.py from oracle/_ref, and the three full table fixtures."""
import json
import os

import numpy as np
import pytest

from sdr_receiver_dvb_t2_b200 import engine as E
from tools.make_golden_pilots import combos, digest, key
from tools.make_golden_tables import CONFIGS, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_mode_equals_the_reference_digest():
    g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'pilot_digests.json')))
    cs = combos()
    assert len(cs) == len(g['digests']) == 364
    bad = []
    for c in cs:
        m = E.mode_init(c[0], c[1], c[2], c[3], g['n_data'], c[4])
        if digest(E.mode_tables(m), with_p2=c[1]) != g['digests'][key(c)]:
            bad.append(key(c))
    assert not bad, bad


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_native_tables_equal_the_reference_fixture(name):
    c = CONFIGS[name]
    want = load(os.path.join(ROOT, 'tests', 'golden', 'tables_%s.npz' % name))
    got = E.mode_tables(E.mode_init(c['fft'], c['ext'], c['pp'], c['gi'], c['n_data']))
    assert got['p'] == want['p']
    for k, v in want.items():
        if k != 'p' and np.size(v):
            assert np.array_equal(np.asarray(v), np.asarray(got[k])), k


def test_undefined_combinations_are_rejected():
    for pp in (1, 3, 5):                               # 32K has no PP1 / PP3 / PP5 (EN 302 755 table 48)
        with pytest.raises(E.T2Error):
            E.mode_init('32K', True, pp, '1/128', 10)
    m = E.Mode()
    assert E.lib().t2b200_mode_init(0, 1, 6, 4, 10, 0, m) == E.ERR_ARG       # 2K: not an N_P2 == 1 mode
