"""K5/K6 parity: CUDA LDPC decoder (through the C-ABI) vs the CPU oracle and the reference's golden
vectors.  Integer work: everything is compared bit-exactly -- hard bits, int8 posteriors, trial
counts -- including non-converged groups."""
import hashlib

import numpy as np
import pytest

from oracle import pyoracle as O
from sdr_receiver_dvb_t2_b200 import engine as E
from tools.make_golden_ldpc import EBN0, FULL

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('code', list(range(12)))
def test_group32_matches_reference_golden(engine, golden_ldpc, code):
    """all twelve PLP codes against vectors made by the unmodified reference (trials 25 and 2)"""
    g = golden_ldpc
    if code in FULL:
        llr = g['c%d_llr' % code]
    else:
        llr, _ = O.make_llr(code, 32, EBN0[code], seed=1000 + code)
        if sha(llr) != str(g['c%d_llr_sha' % code]):
            pytest.skip('numpy generator differs from the one that made the fixtures')
    for t in (25, 2):
        r = engine.ldpc_decode(code, llr, flags=E.LDPC_GROUP32 | E.LDPC_WANT_POST, max_trials=t)
        assert (r['trials_left'] == int(g['c%d_t%d' % (code, t)])).all()
        assert sha(r['bits']) == str(g['c%d_t%d_bits_sha' % (code, t)])
        assert sha(r['post']) == str(g['c%d_t%d_post_sha' % (code, t)])


@pytest.mark.parametrize('code,eb', [(2, 2.5), (3, 3.0), (7, 2.5), (12, 0.8), (13, 1.0), (14, 1.2), (6, 1.0)])
def test_group32_matches_oracle_multi_group(engine, code, eb):
    """several groups + a ragged tail group, compared with the port oracle group by group"""
    n = 32 * 2 + 5
    llr, info = O.make_llr(code, n, eb, seed=31 + code)
    r = engine.ldpc_decode(code, llr, flags=E.LDPC_GROUP32 | E.LDPC_WANT_POST)
    for g0 in range(0, n, 32):
        sl = slice(g0, min(n, g0 + 32))
        tr, bits, post = O.port_ldpc_decode(code, llr[sl], 25, want_post=True)
        assert (r['trials_left'][sl] == tr).all()
        assert (r['iterations'][sl] == 25 - max(tr, -1) - (1 if tr < 0 else 0)).all()
        assert np.array_equal(r['bits'][sl], bits)
        assert np.array_equal(r['post'][sl], post)


def test_native_mode_per_codeword_exit(engine):
    """without GROUP32 every codeword is its own group of one lane"""
    code = 2
    llr, info = O.make_llr(code, 12, 2.6, seed=5)
    r = engine.ldpc_decode(code, llr, flags=E.LDPC_WANT_POST)
    for i in range(12):
        tr, bits, post = O.port_ldpc_decode(code, llr[i:i + 1], 25, want_post=True)
        assert r['trials_left'][i] == tr
        assert np.array_equal(r['bits'][i], bits[0]) and np.array_equal(r['post'][i], post[0])
    assert np.array_equal(r['bits'], info)


def test_fused_bch_strip_descramble_and_packing(engine):
    code = 2
    N, K, KB = engine.ldpc_geometry(code)
    llr, info = O.make_llr(code, 32, 2.8, seed=9)
    ref = O.bch_strip_descramble(info, K, KB)
    r = engine.ldpc_decode(code, llr, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE)
    assert r['bits'].shape == (32, KB) and np.array_equal(r['bits'], ref)
    rp = engine.ldpc_decode(code, llr, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE | E.LDPC_PACK_BITS)
    assert np.array_equal(rp['bits'], np.packbits(ref, axis=1))
    # stand-alone K6 on the byte-per-bit output of a plain decode
    plain = engine.ldpc_decode(code, llr, flags=E.LDPC_GROUP32)
    assert np.array_equal(engine.bch_descramble(code, plain['bits']), ref)


def test_edge_cases(engine):
    code = 7
    N, K, KB = engine.ldpc_geometry(code)
    # empty batch
    r = engine.ldpc_decode(code, np.zeros((0, N), np.int8))
    assert r['bits'].shape == (0, K)
    # all-zero LLRs: every check "fails" (zero posterior), the group runs out of trials
    z = np.zeros((3, N), np.int8)
    r = engine.ldpc_decode(code, z, flags=E.LDPC_GROUP32 | E.LDPC_WANT_POST)
    tr, bits, post = O.port_ldpc_decode(code, z, 25, want_post=True)
    assert tr == -1 and (r['trials_left'] == -1).all() and np.array_equal(r['post'], post)
    # saturated inputs (+-127/-128) exercise the int8 saturation paths
    rng = np.random.default_rng(3)
    s = rng.choice(np.array([-128, -127, 127, 0, 1, -1], np.int8), size=(32, N))
    r = engine.ldpc_decode(code, s, flags=E.LDPC_GROUP32 | E.LDPC_WANT_POST, max_trials=4)
    tr, bits, post = O.port_ldpc_decode(code, s, 4, want_post=True)
    assert (r['trials_left'] == tr).all() and np.array_equal(r['post'], post)
    with pytest.raises(Exception):
        engine.ldpc_decode(99, np.zeros((1, N), np.int8))


def test_device_pointers_and_large_batch_properties(engine):
    """BASELINE-size batch with inputs resident in HBM: size-independent properties only --
    every converged word re-encodes to itself (is a codeword), equals the transmitted info."""
    import torch
    code = 2
    n = 32 * 40
    llr, info = O.make_llr(code, n, 2.9, seed=123)
    d = torch.from_numpy(llr).cuda()
    r = engine.ldpc_decode(code, d, flags=E.LDPC_GROUP32)
    engine.sync()
    bits = r['bits'].cpu().numpy()
    tl = r['trials_left'].cpu().numpy()
    assert (tl >= 0).all()
    assert np.array_equal(bits, info)


def test_host_buffer_pipeline_equals_device_path(engine):
    """more than 512 codewords in HOST memory take the chunked H2D | decode | D2H pipeline (two chunks here, the second
    ragged): same bits, trial counts and iteration counts as one launch on device-resident buffers"""
    import torch
    code = 7
    n = 512 + 3 * 32 + 9
    llr, info = O.make_llr(code, n, 2.6, seed=41)
    flags = E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE
    h = engine.ldpc_decode(code, llr, flags=flags)                       # numpy in, numpy out: pipelined path
    d = engine.ldpc_decode(code, torch.from_numpy(llr).cuda(), flags=flags)
    engine.sync()
    # lock-step groups are formed inside each 512-codeword chunk, i.e. at the same codeword boundaries as in one launch
    assert np.array_equal(h['bits'], d['bits'].cpu().numpy())
    assert np.array_equal(h['trials_left'], d['trials_left'].cpu().numpy())
    assert np.array_equal(h['iterations'], d['iterations'].cpu().numpy())
    N, K, KB = engine.ldpc_geometry(code)
    assert np.array_equal(h['bits'][h['trials_left'] >= 0], O.bch_strip_descramble(info, K, KB)[h['trials_left'] >= 0])
