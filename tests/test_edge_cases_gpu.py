"""Edge cases of every C-ABI entry point on the GPU: empty batches are no-ops, bad arguments and unconfigured stages
return error codes (the binding raises) instead of touching memory, and a failed call leaves the context usable."""
import ctypes as C

import numpy as np
import pytest

from sdr_receiver_dvb_t2_b200 import engine as E

pytestmark = pytest.mark.gpu


def test_empty_batches_are_no_ops(engine):
    assert engine.fft(np.zeros((0, 4096), np.complex64)).shape == (0, 4096)
    assert engine.bch_descramble(2, np.zeros((0, 43200), np.uint8)).shape == (0, 43040)
    engine.ti_configure(7, 0, 2, 4)
    assert engine.ti_deinterleave(7, np.zeros(0, np.complex64), []).shape == (0,)
    d = engine.demap(np.zeros(0, np.complex64), [], 2, 1, 0, 1)
    assert d['llr'].shape == (0, 16200)
    ts, dl, st = engine.ts_packetize(np.zeros((0, 9552), np.uint8))
    assert len(ts) == 0 and len(dl) == 0


def test_bad_arguments_are_rejected(engine):
    L, h = engine.L, engine.h
    x = np.zeros((1, 5000), np.complex64)
    with pytest.raises(E.T2Error):
        engine.fft(x)                                                    # not a power of two
    with pytest.raises(E.T2Error):
        engine.fft(np.zeros((1, 128), np.complex64))                     # below 256
    assert L.t2b200_fft(h, 4096, None, 1, None) == E.ERR_ARG
    assert L.t2b200_ldpc_decode(h, 2, None, 1, None, None, None, None, 25, 0) == E.ERR_ARG
    llr = np.zeros((1, 64800), np.int8)
    assert L.t2b200_ldpc_decode(h, 2, llr.ctypes.data, 1, None, None, None, None, 0, 0) == E.ERR_ARG      # max_trials
    assert L.t2b200_ldpc_decode(h, 2, llr.ctypes.data, 1, None, None, None, None, 25, E.LDPC_WANT_POST) == E.ERR_ARG
    assert L.t2b200_ldpc_decode(h, 12, llr.ctypes.data, 1, None, None, None, None, 25, E.LDPC_BCH_DESCRAMBLE) == E.ERR_ARG  # L1 code: no BCH geometry
    with pytest.raises(E.T2Error):
        engine.ti_deinterleave(200, np.zeros(2700, np.complex64), [1])   # PLP never configured
    engine.ti_configure(8, 0, 2, 2)
    with pytest.raises(E.T2Error):
        engine.ti_deinterleave(8, np.zeros(3 * 2700, np.complex64), [3])  # more FEC blocks than configured
    with pytest.raises(E.T2Error):
        engine.demap(np.zeros(2700, np.complex64), [1], 5, 0, 0, 0)      # unknown modulation
    with pytest.raises(E.T2Error):
        engine.ts_packetize(np.zeros((1, 100), np.uint8))                # shorter than a header + one packet
    assert L.t2b200_set_option(h, 99, 1) == E.ERR_ARG
    assert b'unknown option' in L.t2b200_last_error(h)


def test_context_survives_errors(engine):
    """after the rejected calls above the engine still decodes"""
    from oracle import pyoracle as O
    llr, info = O.make_llr(7, 32, 3.0, seed=77)
    r = engine.ldpc_decode(7, llr, flags=E.LDPC_GROUP32)
    assert np.array_equal(r['bits'], info)


def test_single_symbol_live_mode_calls(engine):
    """B = 1 calls (the per-symbol live path of the facade) agree with the batched ones"""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((3, 16384)) + 1j * rng.standard_normal((3, 16384))).astype(np.complex64)
    whole = engine.fft(x)
    for i in range(3):
        assert np.array_equal(engine.fft(x[i:i + 1])[0], whole[i])


def test_frame_pipeline_rejects_a_geometry_that_differs_from_the_symbol_tables(engine):
    """t2b200_frames_configure cross-checks fft_size / cells per symbol / symbols per frame against what t2b200_eq_configure
    stored (a mismatch would scatter cells across neighbouring symbols), and t2b200_frames_decode bounds max_trials"""
    m = E.mode_init('16K', True, 7, '1/128', 10)
    assert engine.L.t2b200_eq_configure_mode(engine.h, C.byref(m)) == E.OK
    engine.ti_configure(0, 0, 2, 16)
    good = dict(fft_size=m.fft_size, len_frame=m.len_frame, n_p2=1, l_fc=0, c_p2=m.c_p2, c_data=m.c_data, n_fc=0, first_cell=2200,
                plp=0, mod=2, rotation=1, fec_type=0, code_rate=1, n_blocks=32, ti_len=2)
    engine.frames_configure(**good)
    for k, v in (('c_data', m.c_data - 2), ('c_p2', m.c_p2 + 6), ('fft_size', 32768), ('len_frame', m.len_frame + 3)):
        with pytest.raises(E.T2Error):
            engine.frames_configure(**dict(good, **{k: v}))
    engine.frames_configure(**good)
    iq = np.zeros((1, m.len_frame, m.fft_size), np.complex64)
    out = np.zeros((32, 9552), np.uint8)
    assert engine.L.t2b200_frames_decode(engine.h, iq.ctypes.data, 1, out.ctypes.data, None, None, None, None, 64, 3) == E.ERR_ARG
