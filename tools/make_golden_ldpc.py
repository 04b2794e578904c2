#!/usr/bin/env python3
"""Generate tests/golden/ldpc_ref.npz from the UNMODIFIED reference decoder (oracle/_ref).

Runs only in the build container (needs /root/reference to have been compiled by oracle/Makefile).
For each of the twelve PLP codes: 32 seeded noisy codewords (oracle.pyoracle.make_llr), decoded by
the reference with trials = 25 and with trials = 2 (a fixed-iteration snapshot that exercises the
posterior arithmetic before convergence).  Large arrays are stored as SHA-256 digests (inputs are
re-generated from the seed and checked against their digest first); two short codes keep the full
vectors so the fixtures do not depend on the generator at all.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import pyoracle as O  # noqa: E402

# Eb/N0 per code id chosen near the waterfall so that several iterations are needed
EBN0 = {0: 1.2, 1: 2.3, 2: 2.6, 3: 3.1, 4: 3.6, 5: 4.0, 6: 1.2, 7: 2.6, 8: 3.0, 9: 3.5, 10: 3.9, 11: 4.4}
FULL = (7, 11)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build()
    out = {}
    for code in range(12):
        llr, info = O.make_llr(code, 32, EBN0[code], seed=1000 + code)
        r25, b25, p25 = O.ref_ldpc_decode32(code, llr, 25, want_post=True)
        r2, b2, p2 = O.ref_ldpc_decode32(code, llr, 2, want_post=True)
        out['c%d_llr_sha' % code] = sha(llr)
        out['c%d_t25' % code] = np.int32(r25)
        out['c%d_t25_bits_sha' % code] = sha(b25)
        out['c%d_t25_post_sha' % code] = sha(p25)
        out['c%d_t2' % code] = np.int32(r2)
        out['c%d_t2_bits_sha' % code] = sha(b2)
        out['c%d_t2_post_sha' % code] = sha(p2)
        out['c%d_biterr' % code] = np.int32((b25 != info).sum())
        if code in FULL:
            out['c%d_llr' % code] = llr
            out['c%d_t25_bits' % code] = np.packbits(b25, axis=1)
            out['c%d_t25_post' % code] = p25
            out['c%d_t2_post' % code] = p2
        print(code, 'trials left', r25, r2, 'bit errors', int((b25 != info).sum()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'ldpc_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
