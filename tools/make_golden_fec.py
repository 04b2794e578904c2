#!/usr/bin/env python3
"""Generate tests/golden/fec_ref.npz: TI de-interleaver + soft demapper vectors produced by the
UNMODIFIED reference (oracle/_ref/libref_chain.so: time_deinterleaver -> llr_demapper), one process
per configuration because the reference stages keep static state.  Small configurations keep full
vectors; the BASELINE-size ones keep SHA-256 digests of inputs (re-generated from the seed) and outputs.
"""
import hashlib
import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

NORM = [0.707106781, 0.316227766, 0.15430335, 0.076696499]
ROT = [0.506145483, 0.293215314, 0.150098316, 0.062418810]

# name: mod, cod, rot, fec(1 normal), blocks per frame, ti_len, frames, C/N dB, keep full vectors
# The reference's 16-QAM and 256-QAM demappers step 4 cells at a time and only notice the end of a FECFRAME when
# the bit count hits it exactly (llr_demapper.cpp:339,742): with 16 200-bit frames (4050 / 2025 cells) that never
# happens and the loop runs off its tables -- short frames work in the reference only for QPSK and 64-QAM.
CONFIGS = {
    'A_s64_r35': dict(mod=2, cod=1, rot=1, fec=0, nb=8, ti_len=1, frames=4, cn=16.0, full=True),
    'D_sqpsk_r12': dict(mod=0, cod=0, rot=1, fec=0, nb=4, ti_len=1, frames=8, cn=5.0, full=True),
    'G_s64_r23': dict(mod=2, cod=2, rot=0, fec=0, nb=11, ti_len=3, frames=3, cn=16.0, full=False),
    'B_n16_r12': dict(mod=1, cod=0, rot=0, fec=1, nb=4, ti_len=2, frames=8, cn=11.0, full=False),
    'E_n256_r23': dict(mod=3, cod=2, rot=1, fec=1, nb=202, ti_len=3, frames=1, cn=19.0, full=False),
    'F_n64_r35': dict(mod=2, cod=1, rot=1, fec=1, nb=33, ti_len=1, frames=1, cn=15.0, full=False),
    'H_n256_r34': dict(mod=3, cod=3, rot=1, fec=1, nb=40, ti_len=2, frames=1, cn=20.0, full=False),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ti_split(nb, ti_len):
    """FEC blocks per TI block (time_deinterleaver.cpp:275-282)"""
    base = nb // ti_len
    return [base + (1 if j >= ti_len - nb % ti_len else 0) for j in range(ti_len)]


def synth_cells(cfg, n, rng):
    """rotated-QAM cells + AWGN (any complex input exercises the path; QAM keeps the statistics realistic)"""
    m = 1 << (cfg['mod'] + 1)
    a = NORM[cfg['mod']]
    c = ((2 * rng.integers(0, m, n) - (m - 1)) + 1j * (2 * rng.integers(0, m, n) - (m - 1))) * a
    if cfg['rot']:
        c = c * np.exp(1j * ROT[cfg['mod']])
    sig = np.sqrt(10 ** (-cfg['cn'] / 10) / 2)
    return (c + sig * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)


def frame_streams(name, cfg):
    """per frame: (p2 cells, list of data-symbol cell arrays) in the C32 geometry, carrying exactly the PLP's cells
    (no dummy cells: the reference would assemble further TI blocks out of them)"""
    rng = np.random.default_rng(sum(map(ord, name)))
    cpf = (64800 if cfg['fec'] else 16200) // (2 * (cfg['mod'] + 1))
    need = cfg['nb'] * cpf
    out = []
    for _ in range(cfg['frames']):
        n_p2 = min(22432, 2200 + need)
        p2 = synth_cells(cfg, n_p2, rng)
        left = need - (n_p2 - 2200)
        syms = []
        while left > 0:
            syms.append(synth_cells(cfg, min(27404, left), rng))
            left -= len(syms[-1])
        out.append((p2, syms))
    return out


def run_config(name):
    from oracle import pyoracle as O
    cfg = CONFIGS[name]
    rx = O.RefRx('32K', True, 7, '1/128', 59)
    plps = [dict(id=0, cod=cfg['cod'], mod=cfg['mod'], rot=cfg['rot'], fec=cfg['fec'],
                 blocks_max=max(ti_split(cfg['nb'], cfg['ti_len'])), ti_len=cfg['ti_len'], ti_type=0)]
    fec = O.RefFec(rx, plps, l1_post_size=360)
    fec.chain(after_ti=True, after_demap=False)
    cpf = (64800 if cfg['fec'] else 16200) // (2 * (cfg['mod'] + 1))
    streams = []
    for p2, syms in frame_streams(name, cfg):
        fec.feed_p2([0], [cfg['nb']], p2)
        for s in syms:
            fec.feed(s)
        streams.append(np.concatenate([p2[2200:]] + syms)[:cfg['nb'] * cpf])
    t = fec.taps()
    res = {'in': np.concatenate(streams), 'ti': t['ti_cells'], 'llr': t['llr'], 'snr': t['snr'],
           'ti_sizes': t['ti_sizes']}
    return name, res


def main():
    from oracle import pyoracle as O
    O.build()
    out = {}
    with mp.get_context('spawn').Pool(4, maxtasksperchild=1) as pool:   # fresh process per config (static state)
        jobs = [(n, pool.apply_async(run_config, (n,))) for n in CONFIGS]
        for n, j in jobs:
            name, r = j.get(timeout=120)           # a hang in the reference must not hang the generator
            cfg = CONFIGS[name]
            out[name + '_in_sha'] = sha(r['in'])
            out[name + '_ti_sha'] = sha(r['ti'])
            out[name + '_llr_sha'] = sha(r['llr'])
            out[name + '_snr'] = r['snr']
            out[name + '_ti_sizes'] = r['ti_sizes']
            out[name + '_n_llr'] = np.int64(r['llr'].size)
            if cfg['full']:
                out[name + '_llr'] = r['llr']
            print(name, 'cells', r['in'].size, 'ti blocks', r['ti_sizes'].tolist(), 'llr bytes', r['llr'].size,
                  'snr', r['snr'][:3])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'fec_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
