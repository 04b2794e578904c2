#!/usr/bin/env python3
"""Generate tests/golden/e2e_ref.json: what the UNMODIFIED reference receiver (oracle/_ref/libref_chain.so,
dvbt2_demodulator::execute down to bb_de_header's datagrams) emits for the synthetic int16 I/Q streams of
tests/e2e_helpers.py -- SHA-256 of the transport stream, its length, the number of BBFRAMEs, which transmitted BBFRAME
comes out first -- plus the digest of the I/Q itself so that a test can tell a changed modulator from a changed receiver."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def main():
    from oracle import pyoracle as O
    from tests import e2e_helpers as H
    O.build()
    out = {}
    for name in H.CONFIGS:
        t = H.run_reference(name)
        _, _, bb, _ = H.make_stream(name)
        first = next(j for j in range(len(bb)) if np.array_equal(bb[j], t['bb_bits'][0]))
        n = len(t['bb_bits'])
        assert np.array_equal(bb[first:first + n], t['bb_bits']), 'the reference did not return the transmitted BBFRAMEs'
        fr = H.decoded_frames(t)
        out[name] = dict(iq_sha=t['iq_sha'], ts_sha=H.sha(t['ts']), ts_bytes=int(len(t['ts'])), n_bbframes=int(n),
                         first_bbframe=int(first), n_datagrams=int(len(t['ts_datagrams'])), frames_decoded=int(fr.shape[0]),
                         fft_in_sha=H.sha(fr), params=t['params'], snr_db=[round(float(x), 3) for x in t['snr'][:4]])
        print(name, out[name])
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'e2e_ref.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
