"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracles.

  port  -> oracle/_port/liboracle_port.so   (plain-C restatement, oracle/port/*.c)
  ref   -> oracle/_ref/*.so                 (the unmodified reference, compiled by oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i8p = np.ctypeslib.ndpointer(np.int8, flags='C')
_u8p = np.ctypeslib.ndpointer(np.uint8, flags='C')

# code ids (include/t2b200.h): fec*6 + rate; fec 0 = NORMAL(64800), 1 = SHORT(16200); 12..14 = L1 codes
RATES = ['1/2', '3/5', '2/3', '3/4', '4/5', '5/6']
K_BCH = [32208, 38688, 43040, 48408, 51648, 53840, 7032, 9552, 10632, 11712, 12432, 13152]


def code_id(fec_normal: bool, rate: str) -> int:
    return (0 if fec_normal else 6) + RATES.index(rate)


def build(port=True, ref=True):
    targets = (['port'] if port else []) + (['ref'] if ref else [])
    subprocess.run(['make', '-s', '-C', _HERE] + targets, check=True)


_port = None
_ref_ldpc = None


def port():
    global _port
    if _port is None:
        p = os.path.join(_HERE, '_port', 'liboracle_port.so')
        if not os.path.exists(p):
            build(port=True, ref=False)
        L = C.CDLL(p)
        L.port_ldpc_decode_group.argtypes = [C.c_int, _i8p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.port_ldpc_decode_group.restype = C.c_int
        L.port_ldpc_bad.argtypes = [C.c_int, _i8p]
        L.port_ldpc_encode.argtypes = [C.c_int, _u8p, _u8p]
        L.port_bch_strip_descramble.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.port_bb_prbs.argtypes = [_u8p, C.c_int]
        _port = L
    return _port


def have_ref(name='libref_ldpc.so'):
    return os.path.exists(os.path.join(_HERE, '_ref', name))


def ref_ldpc():
    global _ref_ldpc
    if _ref_ldpc is None:
        L = C.CDLL(os.path.join(_HERE, '_ref', 'libref_ldpc.so'))
        L.ref_ldpc_new.restype = C.c_void_p
        L.ref_ldpc_new.argtypes = [C.c_int]
        L.ref_ldpc_decode32.argtypes = [C.c_void_p, C.c_int, _i8p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_ldpc_decode32.restype = C.c_int
        _ref_ldpc = L
    return _ref_ldpc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def code_nk(code):
    L = port()
    return L.port_ldpc_code_n(code), L.port_ldpc_code_k(code)


def port_ldpc_decode(code, llr, trials=25, want_post=False):
    """llr int8[lanes][N] -> (trials_left, bits uint8[lanes][K], post int8[lanes][N] | None)"""
    N, K = code_nk(code)
    llr = np.ascontiguousarray(llr, np.int8).reshape(-1, N)
    lanes = llr.shape[0]
    bits = np.empty((lanes, K), np.uint8)
    post = np.empty((lanes, N), np.int8) if want_post else None
    r = port().port_ldpc_decode_group(code, llr, lanes, _ptr(bits), _ptr(post), trials)
    return r, bits, post


_ref_decoders = {}


def ref_ldpc_decode32(code, llr, trials=25, want_post=False):
    """The reference itself on exactly 32 codewords (ldpc_decoder.cpp:248-277)."""
    N, K = code_nk(code)
    llr = np.ascontiguousarray(llr, np.int8).reshape(32, N)
    L = ref_ldpc()
    if code not in _ref_decoders:
        _ref_decoders[code] = L.ref_ldpc_new(code)
    bits = np.empty((32, K), np.uint8)
    post = np.empty((32, N), np.int8) if want_post else None
    r = L.ref_ldpc_decode32(_ref_decoders[code], code, llr, _ptr(bits), _ptr(post), trials)
    return r, bits, post


def ldpc_encode(code, info):
    N, K = code_nk(code)
    info = np.ascontiguousarray(info, np.uint8).reshape(-1, K)
    cw = np.empty((info.shape[0], N), np.uint8)
    for i in range(info.shape[0]):
        port().port_ldpc_encode(code, info[i], cw[i])
    return cw


def bch_strip_descramble(bits, k_ldpc, k_bch):
    bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, k_ldpc)
    out = np.empty((bits.shape[0], k_bch), np.uint8)
    port().port_bch_strip_descramble(bits, bits.shape[0], k_ldpc, k_bch, out)
    return out


def make_llr(code, n_cw, ebn0_db, seed, scale=2.0, all_zero=False):
    """Synthetic decoder input: random info bits, LDPC-encoded, BPSK + AWGN,
    int8 = clip(round(scale*llr)).  Returns (llr int8[n][N], info uint8[n][K])."""
    N, K = code_nk(code)
    rng = np.random.default_rng(seed)
    info = np.zeros((n_cw, K), np.uint8) if all_zero else rng.integers(0, 2, (n_cw, K), dtype=np.uint8)
    cw = ldpc_encode(code, info)
    rate = K / N
    sigma = np.sqrt(1.0 / (2.0 * rate * 10 ** (ebn0_db / 10.0)))
    x = 1.0 - 2.0 * cw.astype(np.float32)
    y = x + sigma * rng.standard_normal(x.shape, dtype=np.float32)
    llr = 2.0 * y / (sigma * sigma)
    return np.clip(np.rint(scale * llr), -128, 127).astype(np.int8), info
