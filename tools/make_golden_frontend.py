#!/usr/bin/env python3
"""Generate tests/golden/frontend_ref.npz: what the UNMODIFIED reference receiver computes in its front-end
(dvbt2_demodulator::execute, DSP/interpolator_farrow.hh, DSP/filter_decimator.h) for a window of chunks of the synthetic 16K
stream -- see tests/fe_helpers.py."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def main():
    from oracle import pyoracle as O
    from tests import fe_helpers as F
    O.build()
    t = F.run_reference()
    F.save_golden(t)
    w = F.load_golden()
    print('chunks', len(w['info']), 'input offset', w['in_offset'], 'file', os.path.getsize(F.GOLDEN), 'bytes')
    print('resampler counts != 2 x input:', [(int(r[0]), int(r[1])) for r in w['info'] if r[1] != 2 * r[0]])


if __name__ == '__main__':
    main()
