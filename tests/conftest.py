import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def engine():
    """One t2b200 context on cuda:0.  Fails (not skips) when the CUDA library cannot run: a silent
    fallback would void every parity claim."""
    import sdr_receiver_dvb_t2_b200 as t2
    eng = t2.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope='session')
def golden_ldpc():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'ldpc_ref.npz'))
