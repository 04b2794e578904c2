"""CPU-oracle receive chain (port functions only) over modulator output: the checker for the chain tests."""
import numpy as np

from oracle import pyoracle as O


def port_receive(t, m, time, max_trials=25, ldpc=True, precision_override=None):
    """t: tables, m: tools.modulator.Modulator (geometry), time complex64 [len_frame][fft_size] of ONE frame.
    -> dict(stream, ti, llr, snr, precision, bits (descrambled K_bch, per 32-group), trials [per group])"""
    p = t['p']
    cells = []
    sro, ph = [], []
    n_data_sym = p['len_frame'] - p['n_p2'] - p['l_fc']
    for i in range(p['len_frame']):
        f = O.port_fft(time[i])
        if i < p['n_p2']:
            c, s, h = O.port_equalize(0, f, p['l_nulls'], p['k_total'], t['p2_map'], t['p2_ref'], t['h_odd_p2'], p['c_p2'], t['amp_p2'])
        elif i < p['n_p2'] + n_data_sym:
            hh = t['h_odd_data'] if i % 2 == 0 else t['h_even_data']
            c, s, h = O.port_equalize(1, f, p['l_nulls'], p['k_total'], t['data_map'][i - p['n_p2']], t['data_ref'][i - p['n_p2']],
                                      hh, p['c_data'], t['amp_sp'], t['amp_cp'])
        else:
            hh = t['h_odd_fc'] if i % 2 == 0 else t['h_even_fc']
            c, s, h = O.port_equalize(2, f, p['l_nulls'], p['k_total'], t['fc_map'], t['fc_ref'], hh, p['n_fc'], t['amp_sp'])
        cells.append(c), sro.append(s), ph.append(h)
    allc = np.concatenate(cells)
    stream = allc[m.p2_start:m.p2_start + m.nb * m.cpf]
    perm = O.port_cell_permutation(max(m.blocks), m.cpf)
    ti = O.port_ti_blocks(stream, m.blocks, m.cpf, perm, [0, 0.0])
    llrs, snrs, precs, off = [], [], [], 0
    for nf in m.blocks:
        llr, snr, pr, _ = O.port_demap(ti[off:off + nf * m.cpf], m.mod, int(m.rot), m.fec, m.cod)
        llrs.append(llr), snrs.append(snr), precs.append(pr)
        off += nf * m.cpf
    llr = np.concatenate(llrs)
    out = {'stream': stream, 'ti': ti, 'llr': llr, 'snr': np.array(snrs, np.float32), 'precision': np.array(precs, np.float32),
           'sro': np.array(sro, np.float32), 'phase': np.array(ph, np.float32)}
    if ldpc:
        N, K = O.code_nk(m.code)
        kb = O.K_BCH[m.code]
        bits, trials = [], []
        for g0 in range(0, (len(llr) // 32) * 32, 32):
            tr, b, _ = O.port_ldpc_decode(m.code, llr[g0:g0 + 32], max_trials)
            bits.append(O.bch_strip_descramble(b, K, kb)), trials.append(tr)
        out['bits'] = np.concatenate(bits) if bits else np.zeros((0, kb), np.uint8)
        out['trials'] = trials
    return out
