"""CPU-side checks of the C-ABI library: it builds, loads, and exports every symbol that
include/t2b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import sdr_receiver_dvb_t2_b200 as t2
from sdr_receiver_dvb_t2_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 't2b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(t2b200_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    L = t2.lib()
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(L, n), n
    assert sorted(engine.SYMBOLS) == names


def test_code_geometry_matches_reference_tables():
    L = t2.lib()
    # ldpc_decoder.cpp:177-245 and bch_decoder.cpp:79-131
    want = {(1, 0): (64800, 32400, 32208), (1, 1): (64800, 38880, 38688), (1, 2): (64800, 43200, 43040),
            (1, 3): (64800, 48600, 48408), (1, 4): (64800, 51840, 51648), (1, 5): (64800, 54000, 53840),
            (0, 0): (16200, 7200, 7032), (0, 1): (16200, 9720, 9552), (0, 2): (16200, 10800, 10632),
            (0, 3): (16200, 11880, 11712), (0, 4): (16200, 12600, 12432), (0, 5): (16200, 13320, 13152)}
    for (fec, rate), (n, k, kb) in want.items():
        c = L.t2b200_ldpc_code_id(fec, rate)
        assert (L.t2b200_ldpc_n(c), L.t2b200_ldpc_k(c), L.t2b200_ldpc_k_bch(c)) == (n, k, kb)
    assert L.t2b200_ldpc_code_id(2, 0) == -1 and L.t2b200_ldpc_code_id(1, 6) == -1


def test_no_cpu_fallback_without_gpu():
    """without a CUDA device the context cannot be created and the binding raises"""
    import torch
    if torch.cuda.is_available():
        return
    L = t2.lib()
    h = ctypes.c_void_p()
    assert L.t2b200_create(0, ctypes.byref(h)) == engine.ERR_CUDA
    try:
        t2.Engine(0)
        assert False, 'Engine() must fail without a GPU'
    except t2.T2Error:
        pass


def test_product_does_not_touch_oracle():
    """nothing under the package may import / link / load the oracle (comments may mention that it exists)"""
    pkg = os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200')
    banned = ('import oracle', 'from oracle', 'pyoracle', 'liboracle', 'libref_', 'oracle/', '_ref/', 'ldpc_port', 'fec_port',
              'eq_port', 'port_ldpc', 'port_demap', 'port_equalize')
    for dp, _, fs in os.walk(pkg):
        if 'build' in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith(('.py', '.cu', '.cpp', '.h', '.cuh', '.hpp', '.inc')):
                s = open(os.path.join(dp, f), errors='ignore').read()
                for b in banned:
                    assert b not in s, (os.path.join(dp, f), b)
    # and the shared library links only CUDA / system libraries
    import subprocess
    out = subprocess.run(['ldd', t2.lib_path()], capture_output=True, text=True).stdout
    assert 'oracle' not in out and 'fftw' not in out
