"""Pin the CPU oracle (oracle/port/ldpc_port.c) before anything trusts it:
 (1) against golden vectors produced by the unmodified reference decoder (tests/golden/ldpc_ref.npz,
     generator tools/make_golden_ldpc.py) -- runs everywhere;
 (2) against the reference itself (oracle/_ref/libref_ldpc.so) on fresh seeds -- runs where that
     library exists (build container; it also travels to the GPU box).
The reference ships no tests or vectors of its own."""
import hashlib

import numpy as np
import pytest

from oracle import pyoracle as O
from tools.make_golden_ldpc import EBN0, FULL


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('code', [7, 11, 9])
def test_port_matches_golden_short(code, golden_ldpc):
    g = golden_ldpc
    if code in FULL:
        llr = g['c%d_llr' % code]
    else:
        llr, _ = O.make_llr(code, 32, EBN0[code], seed=1000 + code)
        if sha(llr) != str(g['c%d_llr_sha' % code]):
            pytest.skip('numpy generator differs from the one that made the fixtures')
    for t in (25, 2):
        r, bits, post = O.port_ldpc_decode(code, llr, t, want_post=True)
        assert r == int(g['c%d_t%d' % (code, t)])
        assert sha(bits) == str(g['c%d_t%d_bits_sha' % (code, t)])
        assert sha(post) == str(g['c%d_t%d_post_sha' % (code, t)])
    if code in FULL:
        r, bits, post = O.port_ldpc_decode(code, llr, 25, want_post=True)
        assert np.array_equal(np.packbits(bits, axis=1), g['c%d_t25_bits' % code])
        assert np.array_equal(post, g['c%d_t25_post' % code])


def test_port_matches_golden_normal_r23(golden_ldpc):
    g = golden_ldpc
    code = 2
    llr, _ = O.make_llr(code, 32, EBN0[code], seed=1000 + code)
    if sha(llr) != str(g['c2_llr_sha']):
        pytest.skip('numpy generator differs from the one that made the fixtures')
    r, bits, post = O.port_ldpc_decode(code, llr, 2, want_post=True)
    assert r == int(g['c2_t2']) and sha(post) == str(g['c2_t2_post_sha'])


def test_nonconverging_group_matches_golden(golden_ldpc):
    """code 6 at 1.2 dB never converges: the reference returns trials < 0 (batch dropped)."""
    g = golden_ldpc
    llr, _ = O.make_llr(6, 32, EBN0[6], seed=1006)
    if sha(llr) != str(g['c6_llr_sha']):
        pytest.skip('numpy generator differs')
    r, bits, post = O.port_ldpc_decode(6, llr, 25, want_post=True)
    assert r == -1 == int(g['c6_t25'])
    assert sha(post) == str(g['c6_t25_post_sha'])


@pytest.mark.skipif(not O.have_ref(), reason='oracle/_ref/libref_ldpc.so not built (no /root/reference)')
@pytest.mark.parametrize('code', [6, 8, 10, 12, 13, 14])
def test_port_matches_reference_fresh_seed(code):
    eb = {6: 1.6, 8: 3.0, 10: 3.9, 12: 0.8, 13: 1.0, 14: 1.2}[code]
    llr, info = O.make_llr(code, 32, eb, seed=77 + code)
    for t in (25, 3):
        r1, b1, p1 = O.ref_ldpc_decode32(code, llr, t, want_post=True)
        r2, b2, p2 = O.port_ldpc_decode(code, llr, t, want_post=True)
        assert r1 == r2
        assert np.array_equal(b1, b2)
        assert np.array_equal(p1, p2)


def test_encoder_produces_codewords():
    """the test encoder's words satisfy every parity check of the decoder (noise-free LLRs pass bad())"""
    for code in (2, 7, 6):
        N, K = O.code_nk(code)
        info = np.random.default_rng(5).integers(0, 2, (2, K), dtype=np.uint8)
        cw = O.ldpc_encode(code, info)
        llr = (20 * (1 - 2 * cw.astype(np.int16))).astype(np.int8)
        for row in llr:
            assert O.port().port_ldpc_bad(code, np.ascontiguousarray(row)) == 0
        r, bits, _ = O.port_ldpc_decode(code, llr, 25)
        assert r == 25 and np.array_equal(bits, info)


def test_bb_prbs_known_prefix():
    """BB scrambler 1+x^14+x^15, init 100101010000000 (EN 302 755 5.2.4): first bits 0000 0011 ..."""
    p = np.zeros(16, np.uint8)
    O.port().port_bb_prbs(p, 16)
    assert p[:8].tolist() == [0, 0, 0, 0, 0, 0, 1, 1]
