#!/usr/bin/env python3
"""Generate tests/golden/pilot_digests.json: SHA-256 digests of the init-time tables (mode parameters, carrier maps,
pilot references, amplitudes) the UNMODIFIED reference builds (oracle/_ref/libref_chain.so: dvbt2_*_parameters_init,
pilot_generator::p2_generator / data_generator) for every SISO 16K / 32K mode EN 302 755 defines: carrier mode x PP1-PP8
x guard interval x PAPR off / tone reservation.  tests/test_pilot_tables.py holds the native builder
(t2b200_mode_init / t2b200_pilot_tables) to them."""
import hashlib
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
GIS = ['1/32', '1/16', '1/8', '1/4', '1/128', '19/128', '19/256']
N_DATA = 19            # > the longest scattered-pilot period (dy = 16)


def digest(t, with_p2=True):
    """t: tables dict (RefRx.tables() layout) -> hex digest; shared by the generator and the test.
    with_p2=False leaves the P2 tables out: the reference's p2_symbol::init always builds them for extended carriers
    (dvbt2_definition.cpp:89), so in normal-carrier modes only the data / frame-closing tables can be compared."""
    h = hashlib.sha256()
    h.update(json.dumps([t['p'][k] for k in sorted(t['p'])]).encode())
    h.update(np.array([t['amp_p2'], t['amp_sp'], t['amp_cp']], np.float32).tobytes())
    for k in ('p2_map', 'p2_ref', 'data_map', 'data_ref', 'fc_map', 'fc_ref'):
        if k in t and (with_p2 or not k.startswith('p2')):
            a = np.ascontiguousarray(t[k], np.int32 if k.endswith('map') else np.float32)
            h.update(k.encode() + a.tobytes())
    return h.hexdigest()


def combos():
    from sdr_receiver_dvb_t2_b200 import engine as E
    out = []
    for fft in ('16K', '32K'):
        for ext in (False, True):
            for pp in range(1, 9):
                for gi in GIS:
                    for papr in (0, 2):
                        try:
                            E.mode_init(fft, ext, pp, gi, N_DATA, papr)
                        except E.T2Error:
                            continue
                        out.append((fft, ext, pp, gi, papr))
    return out


def key(c):
    return '%s_%s_pp%d_gi%s_papr%d' % (c[0], 'ext' if c[1] else 'nrm', c[2], c[3].replace('/', '-'), c[4])


def run(c):
    from oracle import pyoracle as O
    fft, ext, pp, gi, papr = c
    rx = O.RefRx(fft, ext, pp, gi, N_DATA, papr)
    t = rx.tables()
    t['p'] = rx.p
    return key(c), digest(t, with_p2=ext)


def main():
    from oracle import pyoracle as O
    O.build()
    cs = combos()
    out = {}
    with mp.get_context('spawn').Pool(8, maxtasksperchild=1) as pool:
        for k, d in pool.imap_unordered(run, cs):
            out[k] = d
    path = os.path.join(ROOT, 'tests', 'golden', 'pilot_digests.json')
    json.dump({'n_data': N_DATA, 'digests': dict(sorted(out.items()))}, open(path, 'w'), indent=0)
    print(len(out), 'modes ->', path)


if __name__ == '__main__':
    main()
