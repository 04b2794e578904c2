"""K2 parity through the C-ABI: CUDA equaliser vs the CPU oracle, BIT-EXACT on every cell and on both feedback
floats (the kernel repeats the oracle's float operations in the same order with round-to-nearest intrinsics;
the oracle itself is pinned to the reference within its -Ofast tolerance in tests/test_oracle_eq.py)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests.eq_helpers import kind_tables, synth_symbol, tables

pytestmark = pytest.mark.gpu


def run_kind(engine, t, kind, idx_list, rng, snr_db=25.0):
    p = t['p']
    maps, refs, he, ho, n_out, first, amp_main, amp_cp = kind_tables(t, kind)
    engine.eq_configure(kind, first, p['fft_size'], p['k_total'], p['l_nulls'], n_out, maps, refs, he, ho, amp_main, amp_cp)
    freq = np.stack([synth_symbol(t, maps[min(max(i - first, 0), len(maps) - 1)], refs[min(max(i - first, 0), len(maps) - 1)],
                                  rng, snr_db) for i in idx_list])
    cells, sro, ph = engine.equalize(kind, idx_list, freq)
    for s, idx in enumerate(idx_list):
        r = min(max(idx - first, 0), len(maps) - 1)
        h = ho if idx % 2 == 0 else he
        want, wsro, wph = O.port_equalize(kind, freq[s], p['l_nulls'], p['k_total'], maps[r], refs[r], h, n_out, amp_main, amp_cp)
        assert np.array_equal(cells[s].view(np.float32), want.view(np.float32)), (kind, idx)
        assert sro[s] == np.float32(wsro) and ph[s] == np.float32(wph), (kind, idx, sro[s], wsro, ph[s], wph)
    return freq, cells


def test_data_symbols_c32_bit_exact(engine):
    t = tables('c32')
    run_kind(engine, t, 1, [1, 2, 3, 4, 5, 30, 59], np.random.default_rng(11))


def test_p2_symbol_c32_and_c16_bit_exact(engine):
    run_kind(engine, tables('c32'), 0, [0], np.random.default_rng(12))
    run_kind(engine, tables('c16'), 0, [0], np.random.default_rng(13))


def test_data_symbols_c16_bit_exact(engine):
    run_kind(engine, tables('c16'), 1, [1, 2, 7, 8], np.random.default_rng(14))


def test_frame_closing_and_pp4_bit_exact(engine):
    t = tables('c32fc')
    run_kind(engine, t, 1, [1, 2, 3, 4, 11], np.random.default_rng(15))
    run_kind(engine, t, 2, [t['p']['len_frame'] - 1], np.random.default_rng(16))


def test_low_snr_phase_wrap_and_device_buffers(engine):
    """noisy pilots make neighbouring angle estimates straddle +-pi: exercises the reference's asymmetric unwrap;
    inputs / outputs resident on the device"""
    import torch
    t = tables('c32')
    p = t['p']
    maps, refs, he, ho, n_out, first, amp_main, amp_cp = kind_tables(t, 1)
    engine.eq_configure(1, first, p['fft_size'], p['k_total'], p['l_nulls'], n_out, maps, refs, he, ho, amp_main, amp_cp)
    rng = np.random.default_rng(17)
    idx = [1, 2, 3, 4, 5, 6, 7, 8]
    freq = np.stack([synth_symbol(t, maps[i - 1], refs[i - 1], rng, snr_db=-3.0) for i in idx])
    d = torch.from_numpy(freq).cuda()
    cells, sro, ph = engine.equalize(1, idx, d)
    engine.sync()
    sro, ph = sro.cpu().numpy(), ph.cpu().numpy()
    cells = cells.cpu().numpy()
    for s, i in enumerate(idx):
        want, wsro, wph = O.port_equalize(1, freq[s], p['l_nulls'], p['k_total'], maps[i - 1], refs[i - 1],
                                          ho if i % 2 == 0 else he, n_out, amp_main, amp_cp)
        assert np.array_equal(cells[s].view(np.float32), want.view(np.float32))
        assert sro[s] == np.float32(wsro) and ph[s] == np.float32(wph)


def test_eq_errors(engine):
    with pytest.raises(Exception):
        e2 = type(engine)(0)
        e2._eq_nout = {1: (10, 1024)}
        e2.equalize(1, [1], np.zeros((1, 1024), np.complex64))          # tables never configured
