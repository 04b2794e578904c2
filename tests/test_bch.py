"""N3, opt-in: the true BCH decoder the reference leaves as "TODO BCH decode" (bch_decoder.cpp:136).
CPU side: the plain-C decoder (oracle/port/bch_port.c) and the test modulator's encoder agree with each other and with the
code's algebra (every codeword has zero syndromes, up to t errors are corrected, t + 1 are not mis-corrected silently into
the sent word).  GPU side: t2b200_bch_decode equals the C decoder word for word, and the frame pipeline with
T2B200_OPT_BCH_CORRECT returns the transmitted BBFRAMEs."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from tools import modulator as M


def _port():
    L = O.port()
    L.port_bch_encode.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.port_bch_decode.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.port_bch_t.argtypes = [C.c_int]
    return L


def codeword(code, rng):
    """a random BCH codeword of `code` (N_bch = K_ldpc bits, byte per bit) from the modulator's encoder"""
    N, K = O.code_nk(code)
    kb = O.K_BCH[code]
    w = np.zeros(K, np.uint8)
    w[:kb] = rng.integers(0, 2, kb, dtype=np.uint8)
    w[kb:] = M.bch_parity(w[:kb], code >= 6, M.bch_t(code))
    return w


@pytest.mark.parametrize('code', list(range(12)))
def test_encoders_agree_and_codewords_have_zero_syndromes(code):
    L = _port()
    rng = np.random.default_rng(code)
    w = codeword(code, rng)
    c = w.copy()
    L.port_bch_encode(code, c.ctypes.data, len(c))           # the C encoder overwrites the parity bits
    assert np.array_equal(c, w)
    assert L.port_bch_t(code) == M.bch_t(code) and len(w) - O.K_BCH[code] == (14 if code >= 6 else 16) * M.bch_t(code)
    syn = np.ones(24, np.uint16)
    assert L.port_bch_decode(code, c.ctypes.data, len(c), syn.ctypes.data) == 0 and not syn[:2 * M.bch_t(code)].any()


@pytest.mark.parametrize('code', [0, 2, 5, 7, 11])
def test_port_decoder_corrects_up_to_t_errors(code):
    L = _port()
    rng = np.random.default_rng(100 + code)
    w = codeword(code, rng)
    t = M.bch_t(code)
    for ne in (1, 2, t - 1, t):
        e = w.copy()
        e[rng.choice(len(w), ne, replace=False)] ^= 1
        assert L.port_bch_decode(code, e.ctypes.data, len(e), None) == ne and np.array_equal(e, w)
    e = w.copy()
    e[rng.choice(len(w), t + 1, replace=False)] ^= 1
    r = L.port_bch_decode(code, e.ctypes.data, len(e), None)
    assert r == -1 or not np.array_equal(e, w)                # beyond the design distance: flagged (or mis-decoded to ANOTHER codeword)


@pytest.mark.gpu
@pytest.mark.parametrize('code', [0, 2, 3, 5, 6, 7, 11])
def test_gpu_decoder_equals_the_port(engine, code):
    L = _port()
    rng = np.random.default_rng(200 + code)
    t = M.bch_t(code)
    base = codeword(code, rng)
    n_err = [0, 1, 2, t // 2, t - 1, t, t, t + 1, t + 3, 0, 5, t + 1]
    words = np.tile(base, (len(n_err), 1))
    for i, ne in enumerate(n_err):
        if i >= 6:
            words[i] = codeword(code, rng)
        if ne:
            words[i, rng.choice(words.shape[1], ne, replace=False)] ^= 1
    want = words.copy()
    want_cor = np.array([L.port_bch_decode(code, want[i].ctypes.data, want.shape[1], None) for i in range(len(n_err))], np.int32)
    got = words.copy()
    cor = engine.bch_decode(code, got)
    assert np.array_equal(cor, want_cor) and np.array_equal(got, want)
    assert list(cor[:7]) == n_err[:7]
    # device buffers, in place
    import torch
    d = torch.from_numpy(words.copy()).cuda()
    cor = engine.bch_decode(code, d)
    engine.sync()
    assert np.array_equal(cor.cpu().numpy(), want_cor) and np.array_equal(d.cpu().numpy(), want)


@pytest.mark.gpu
def test_frame_pipeline_with_bch_correction(engine):
    """T2B200_OPT_BCH_CORRECT on: frames whose BBFRAMEs carry real BCH parity decode to the transmitted BBFRAMEs (and the
    default path, which strips the parity unread like the reference, gives the same bits)"""
    import torch
    from sdr_receiver_dvb_t2_b200 import engine as E
    from sdr_receiver_dvb_t2_b200.chain import FrameChain
    t = E.mode_tables(E.mode_init('16K', True, 7, '1/128', 30))
    m = M.Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=64, ti_len=2, seed=21, bch=True)
    f = m.frame(noise_cn_db=15.5)
    ch = FrameChain(engine, t, mod=2, cod=1, fec_type=0, n_blocks=64, ti_len=2)
    x = torch.from_numpy(f['time'][None]).cuda()
    plain = ch.decode_frames_fused(x)['bits'].cpu().numpy()
    engine.set_option(E.OPT_BCH_CORRECT, 1)
    try:
        fixed = ch.decode_frames_fused(x)['bits'].cpu().numpy()
    finally:
        engine.set_option(E.OPT_BCH_CORRECT, 0)
    assert np.array_equal(plain, f['bb']) and np.array_equal(fixed, f['bb'])
