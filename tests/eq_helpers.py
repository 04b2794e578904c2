"""Shared helpers for the FFT / equaliser parity tests: table fixtures and synthetic frequency-domain symbols."""
import os

import numpy as np

from tools.make_golden_tables import load as load_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tables(name):
    return load_tables(os.path.join(ROOT, 'tests', 'golden', 'tables_%s.npz' % name))


def synth_symbol(t, cmap, ref, rng, snr_db=25.0, scale=200.0):
    """shifted FFT-domain symbol: pilots = reference * channel, data = 256-QAM * channel, smooth fading + CPE + AWGN"""
    p = t['p']
    k, n = p['k_total'], p['fft_size']
    kk = np.arange(k)
    ch = (1.0 + 0.35 * np.cos(kk / 811.0 + rng.uniform(0, 6))) * np.exp(1j * (rng.uniform(-3, 3) + kk * rng.uniform(-3e-4, 3e-4)))
    x = ((rng.integers(0, 16, k) * 2 - 15) + 1j * (rng.integers(0, 16, k) * 2 - 15)) * 0.076696499
    x = np.where(ref != 0, ref, x)
    x = x * ch + 10 ** (-snr_db / 20) / np.sqrt(2) * (rng.standard_normal(k) + 1j * rng.standard_normal(k))
    f = np.zeros(n, np.complex64)
    f[p['l_nulls']:p['l_nulls'] + k] = (x * scale).astype(np.complex64)
    return f


def kind_tables(t, kind):
    """-> (maps [n][k], refs [n][k], h_even, h_odd, n_out, first_symbol, amp_main, amp_cp)"""
    p = t['p']
    if kind == 0:
        return t['p2_map'][None], t['p2_ref'][None], t['h_even_p2'], t['h_odd_p2'], p['c_p2'], 0, t['amp_p2'], 0.0
    if kind == 1:
        return t['data_map'], t['data_ref'], t['h_even_data'], t['h_odd_data'], p['c_data'], p['n_p2'], t['amp_sp'], t['amp_cp']
    return t['fc_map'][None], t['fc_ref'][None], t['h_even_fc'], t['h_odd_fc'], p['n_fc'], p['len_frame'] - 1, t['amp_sp'], 0.0
