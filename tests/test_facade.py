"""The C++ mirror of the reference's per-stage API (host/t2b200_stages.hpp): it must compile and link against
libt2b200.so without Qt (CPU test), and -- driven symbol by symbol like dvbt2_demodulator::symbol_acquisition --
return the transmitted BBFRAMEs for BASELINE config 4 (GPU test)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import sdr_receiver_dvb_t2_b200 as t2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'cpp', 'facade_check.cpp')
BIN = os.path.join(ROOT, 'tests', 'cpp', 'facade_check')


def build_binary():
    t2.lib()
    libdir = os.path.dirname(t2.lib_path())
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(SRC), os.path.getmtime(t2.lib_path())):
        subprocess.run(['g++', '-std=c++17', '-O1', SRC, '-o', BIN, '-L' + libdir, '-l:libt2b200.so',
                        '-Wl,-rpath,' + libdir], check=True)
    return BIN


def test_facade_compiles_and_links_without_qt():
    b = build_binary()
    assert os.path.exists(b)
    r = subprocess.run([b], capture_output=True)          # no arguments: usage error, but it loads and runs
    assert r.returncode == 2


@pytest.mark.gpu
def test_facade_decodes_config4_frame():
    from tests.eq_helpers import tables
    from tools.modulator import Modulator
    b = BIN if os.path.exists(BIN) else build_binary()
    t = tables('c16')
    p = t['p']
    m = Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=64, ti_len=2, seed=9)
    fr = m.frame(noise_cn_db=15.0)
    with tempfile.TemporaryDirectory() as d:
        fin, fout, fts = os.path.join(d, 'in.bin'), os.path.join(d, 'out.bin'), os.path.join(d, 'ts.bin')
        with open(fin, 'wb') as f:
            hdr = [p['fft_size'], p['k_total'], p['l_nulls'], p['n_p2'], p['n_data'], p['len_frame'], p['l_fc'], p['c_p2'], p['c_data'],
                   p['n_fc'], 0, 1, 2, 1, 0, 32, 2, 0, 360, 64]
            np.array(hdr, np.int32).tofile(f)
            np.array([t['amp_sp'], t['amp_cp'], t['amp_p2']], np.float32).tofile(f)
            t['data_map'].astype(np.int32).tofile(f)
            t['data_ref'].astype(np.float32).tofile(f)
            t['p2_map'].astype(np.int32).tofile(f)
            t['p2_ref'].astype(np.float32).tofile(f)
            for k in ('h_even_data', 'h_odd_data', 'h_even_p2', 'h_odd_p2'):
                t[k].astype(np.int32).tofile(f)
            fr['time'].astype(np.complex64).tofile(f)
        r = subprocess.run([b, fin, fout, fts], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert 'bbframes 64' in r.stdout
        assert 'frontend 4096 level 0.249' in r.stdout         # the front-end mirror: 4096 / 2^14 = 0.25 less the DC the averager has taken out after 3000 samples
        got = np.fromfile(fout, np.uint8).reshape(64, -1)
        ts = np.fromfile(fts, np.uint8)
    assert np.array_equal(got, fr['bb'])
    # ... and the bb_de_header mirror turned them into the TS the reference's bb_de_header makes of them (oracle port)
    from oracle import pyoracle as O
    port = O.PortTs()
    assert 'datagrams 64' in r.stdout
    assert np.array_equal(ts, np.concatenate([port.feed(b) for b in fr['bb']]))
