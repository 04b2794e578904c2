"""End-to-end parity against the reference's OWN receiver: synthetic int16 I/Q with P1 + L1 signalling ->
unmodified dvbt2_demodulator::execute ... bb_de_header (oracle/_ref/libref_chain.so) -> transport stream.

CPU side (this file): (1) the reference run reproduces the committed golden digests (tests/golden/e2e_ref.json); (2) the
oracle port chain, started from the FFT windows the reference cut out of the stream, returns the reference's BBFRAMEs and
byte-identical TS datagrams -- i.e. the restatement the GPU tests are checked against IS the reference end to end.
The GPU side is tests/test_e2e_gpu.py."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import e2e_helpers as H
from tests.chain_helpers import port_receive

pytestmark = pytest.mark.skipif(not O.have_ref('libref_chain.so'), reason='compiled reference (oracle/_ref) not present')

_cache = {}


def reference(name):
    if name not in _cache:
        _cache[name] = H.run_reference(name)
    return _cache[name]


@pytest.mark.parametrize('name', sorted(H.CONFIGS))
def test_reference_receiver_reproduces_the_golden_ts(name):
    g = H.golden()[name]
    t = reference(name)
    assert t['iq_sha'] == g['iq_sha'], 'the test transmitter changed: regenerate tests/golden/e2e_ref.json'
    assert t['params'] == g['params']                      # what the reference read out of P1 + L1-pre / L1-post
    assert len(t['ts']) == g['ts_bytes'] and len(t['bb_bits']) == g['n_bbframes']
    assert H.sha(t['ts']) == g['ts_sha']
    assert H.sha(H.decoded_frames(t)) == g['fft_in_sha']


@pytest.mark.parametrize('name', sorted(H.CONFIGS))
def test_port_chain_from_the_reference_fft_windows_gives_the_reference_ts(name):
    t = reference(name)
    g = H.golden()[name]
    tx = H.transmitter(name)
    frames = H.decoded_frames(t)
    assert frames.shape[0] == g['frames_decoded']
    bb = np.concatenate([port_receive(tx.tables, tx.mod, f)['bits'] for f in frames])
    assert np.array_equal(bb, t['bb_bits'])                # BBFRAMEs of the reference's bch_decoder, bit for bit
    p = O.PortTs()
    ts = np.concatenate([p.feed(b) for b in bb])
    assert np.array_equal(ts, t['ts'])                     # and the datagrams bb_de_header sent
