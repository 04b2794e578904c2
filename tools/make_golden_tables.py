#!/usr/bin/env python3
"""Generate tests/golden/tables_<cfg>.npz: the init-time tables (carrier maps, pilot references,
frequency de-interleaver addresses, amplitudes) that the reference's pilot_generator and
address_freq_deinterleaver build for a transmission mode, dumped from the UNMODIFIED reference
(oracle/_ref/libref_chain.so).  In the drop-in these objects stay the reference's own and the facade hands
their tables to the engine (t2b200_eq_configure); the fixtures let tests and bench.py run on the GPU box,
where the reference is absent.  Stored compactly: unique map rows + per-symbol row index, pilot references
as signed amplitude codes, addresses as uint16."""
import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

CONFIGS = {
    'c32': dict(fft='32K', ext=True, pp=7, gi='1/128', n_data=59),      # configs 1, 2, 5
    'c16': dict(fft='16K', ext=True, pp=7, gi='1/128', n_data=59),      # config 4
    'c32fc': dict(fft='32K', ext=True, pp=4, gi='1/16', n_data=12),     # has a frame-closing symbol
}


def encode_ref(ref, amps):
    code = np.zeros(ref.shape, np.int8)
    for i, a in enumerate(amps, 1):
        code[ref == np.float32(a)] = i
        code[ref == np.float32(-a)] = -i
    assert ((code != 0) == (ref != 0)).all()
    return code


def decode_ref(code, amps):
    out = np.zeros(code.shape, np.float32)
    for i, a in enumerate(amps, 1):
        out[code == i] = np.float32(a)
        out[code == -i] = np.float32(-a)
    return out


def run(name):
    from oracle import pyoracle as O
    c = CONFIGS[name]
    rx = O.RefRx(c['fft'], c['ext'], c['pp'], c['gi'], c['n_data'])
    t = rx.tables()
    p = rx.p
    amps = [t['amp_sp'], t['amp_cp'], t['amp_p2']]
    out = {k: np.int32(v) for k, v in p.items()}
    out['amps'] = np.array(amps, np.float32)
    rows, idx = np.unique(t['data_map'], axis=0, return_inverse=True)
    out['data_map_rows'] = rows.astype(np.int8)
    out['data_map_idx'] = idx.astype(np.int16)
    # pilot references differ per symbol (PN sequence): keep the sign pattern per symbol, bit-packed
    code = encode_ref(t['data_ref'], amps)
    assert np.array_equal(decode_ref(code, amps), t['data_ref'])
    out['data_ref_abs_rows'] = np.abs(code)[np.unique(idx, return_index=True)[1]].astype(np.int8)
    out['data_ref_neg'] = np.packbits(code < 0, axis=1)
    out['p2_map'] = t['p2_map'].astype(np.int8)
    out['p2_ref'] = encode_ref(t['p2_ref'], amps)
    if p['l_fc']:
        out['fc_map'] = t['fc_map'].astype(np.int8)
        out['fc_ref'] = encode_ref(t['fc_ref'], amps)
    for k in ('p2', 'data', 'fc'):
        n = {'p2': p['c_p2'], 'data': p['c_data'], 'fc': p['n_fc']}[k]
        out['h_even_' + k] = t['h_even_' + k][:n].astype(np.uint16)
        out['h_odd_' + k] = t['h_odd_' + k][:n].astype(np.uint16)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'tables_%s.npz' % name)
    np.savez_compressed(path, **out)
    # round trip
    back = load(path)
    for k in ('data_map', 'data_ref', 'p2_map', 'p2_ref'):
        assert np.array_equal(back[k], t[k]), k
    return name, os.path.getsize(path), p


def load(path):
    """fixture -> dict with the same arrays RefRx.tables() gives (+ the mode parameters under 'p')"""
    g = np.load(path)
    amps = [float(x) for x in g['amps']]
    t = {'p': {k: int(g[k]) for k in ('fft_size', 'k_total', 'l_nulls', 'c_p2', 'c_data', 'n_fc', 'c_fc', 'n_data',
                                       'len_frame', 'l_fc', 'n_p2', 'guard_interval_size', 'k_ext')}}
    t['amp_sp'], t['amp_cp'], t['amp_p2'] = amps
    idx = g['data_map_idx'].astype(np.int64)
    t['data_map'] = g['data_map_rows'].astype(np.int32)[idx]
    k = t['p']['k_total']
    neg = np.unpackbits(g['data_ref_neg'], axis=1)[:, :k].astype(bool)
    code = g['data_ref_abs_rows'].astype(np.int8)[idx]
    code = np.where(neg, -code, code).astype(np.int8)
    t['data_ref'] = decode_ref(code, amps)
    t['p2_map'] = g['p2_map'].astype(np.int32)
    t['p2_ref'] = decode_ref(g['p2_ref'], amps)
    if 'fc_map' in g:
        t['fc_map'] = g['fc_map'].astype(np.int32)
        t['fc_ref'] = decode_ref(g['fc_ref'], amps)
    for kk in ('p2', 'data', 'fc'):
        t['h_even_' + kk] = g['h_even_' + kk].astype(np.int32)
        t['h_odd_' + kk] = g['h_odd_' + kk].astype(np.int32)
    return t


def main():
    from oracle import pyoracle as O
    O.build()
    with mp.get_context('spawn').Pool(3, maxtasksperchild=1) as pool:
        for j in [pool.apply_async(run, (n,)) for n in CONFIGS]:
            print(j.get(timeout=120))


if __name__ == '__main__':
    main()
