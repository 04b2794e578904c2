"""The test modulator against the CPU oracle chain (no GPU): a frame built by tools/modulator.py goes through the
port restatements of every reference stage and the BBFRAME bits come back -- i.e. the modulator is the inverse of
the receive path, and the chain of oracles is self-consistent end to end (64-QAM r3/5, the mode the reference can
actually decode on AWGN)."""
import numpy as np

from tests.chain_helpers import port_receive
from tests.eq_helpers import tables
from tools.modulator import Modulator


def test_modulator_roundtrip_through_oracle_chain_c16_64qam():
    t = tables('c16')
    m = Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=32, ti_len=1, seed=4)
    fr = m.frame(noise_cn_db=15.0)
    r = port_receive(t, m, fr['time'])
    assert r['trials'][0] >= 0                                   # the group converged
    assert np.array_equal(r['bits'], fr['bb'][:32])              # BBFRAME bits (what bb_de_header packetises into TS)
    # equalised cells sit on the transmitted ones up to noise
    assert np.abs(r['ti'].reshape(32, -1) - fr['cells']).std() < 0.2
