#!/usr/bin/env python3
"""Generate tests/golden/ts_ref.npz: the datagrams the UNMODIFIED reference bb_de_header (oracle/_ref/libref_chain.so)
emits for the BBFRAME streams of tests/test_oracle_ts.py::CASES (inputs are re-generated from the seed)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import pyoracle as O          # noqa: E402
from tests.test_oracle_ts import CASES, make   # noqa: E402


def main():
    rx = O.RefRx('32K', True, 7, '1/128', 59)
    out = {}
    for case in CASES:
        fec = O.RefFec(rx, [dict(id=0, cod=1, mod=2, rot=1, fec=0, blocks_max=8, ti_len=1, ti_type=0)], 360)
        fec.clear()
        frames, _ = make(case)
        for f in frames:
            fec.deheader(f)
        t = fec.taps()
        out[case + '_ts'] = t['ts'].copy()
        out[case + '_len'] = t['ts_datagrams'].copy()
        print(case, len(frames), 'frames ->', list(t['ts_datagrams']))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'ts_ref.npz'), **out)


if __name__ == '__main__':
    main()
