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
        L.port_cell_permutation.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.port_ti_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.port_demap_address_for.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.port_demap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_float]
        L.port_demap.restype = C.c_float
        L.port_equalize.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_float, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.port_fft_shift.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.port_atan2_approx.argtypes = [C.c_float, C.c_float]
        L.port_atan2_approx.restype = C.c_float
        L.port_sincos_lut.argtypes = [C.c_void_p, C.c_void_p]
        L.port_fe_reset.argtypes = [C.c_void_p]
        L.port_fe_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.port_fe_cp_correlate.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.port_fe_cp_correlate.restype = C.c_float
        L.port_p1_reset.argtypes = [C.c_void_p]
        L.port_p1_correlate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
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


# ---------------------------------------------------------------------------------------------
# the whole reference receiver, stage by stage (oracle/ref_chain.cc -> oracle/_ref/libref_chain.so)
_ref_chain = None
FFT_MODE = {'16K': 4, '32K': 5}                      # dvbt2_definition.h:121-131
GI = {'1/32': 0, '1/16': 1, '1/8': 2, '1/4': 3, '1/128': 4, '19/128': 5, '19/256': 6}
PARAM_NAMES = ['fft_size', 'k_total', 'l_nulls', 'c_p2', 'c_data', 'n_fc', 'c_fc', 'n_data', 'len_frame', 'l_fc',
               'n_p2', 'guard_interval_size', 'k_ext']
_f32p = np.ctypeslib.ndpointer(np.float32, flags='C')
_i32p = np.ctypeslib.ndpointer(np.int32, flags='C')


def ref_chain():
    """One receiver instance per process (the reference keeps static state)."""
    global _ref_chain
    if _ref_chain is None:
        L = C.CDLL(os.path.join(_HERE, '_ref', 'libref_chain.so'))
        L.ref_rx_init.argtypes = [C.c_int] * 6 + [_i32p]
        L.ref_rx_tables_data.argtypes = [C.c_int, _i32p, _f32p]
        L.ref_rx_tables_p2.argtypes = [_i32p, _f32p]
        L.ref_rx_tables_fc.argtypes = [_i32p, _f32p]
        L.ref_rx_tables_h.argtypes = [C.c_int, _i32p, _i32p]
        L.ref_rx_amps.argtypes = [C.POINTER(C.c_float)] * 3
        L.ref_fft.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_data_symbol.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_fc_symbol.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_p2_symbol.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_fec_start.argtypes = [C.c_int, _i32p, C.c_int, C.c_int]
        L.ref_fec_chain.argtypes = [C.c_int] * 4
        L.ref_fec_feed_p2.argtypes = [_i32p, _i32p, C.c_int, C.c_void_p]
        L.ref_fec_feed.argtypes = [C.c_int, C.c_void_p]
        L.ref_demap.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.ref_ldpc_batch.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.ref_bb_deheader.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ref_ti_permutation.argtypes = [C.c_int, _i32p, C.c_int]
        L.ref_demap_address.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
        for n in ('ti_cells', 'ti_sizes', 'llr', 'ldpc_bits', 'bb_bits', 'bb_len', 'snr', 'ts', 'ts_datagrams'):
            f = getattr(L, 'ref_tap_' + n)
            f.argtypes = [C.c_void_p, C.c_longlong]
            f.restype = C.c_longlong
        L.ref_demod_new.argtypes = [C.c_float, C.c_int]
        L.ref_demod_feed.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_demod_params.argtypes = [_i32p]
        L.ref_tap_fft_arm.argtypes = [C.c_int]
        L.ref_tap_frontend_arm.argtypes = [C.c_int, C.c_int]
        L.ref_p1_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for n in ('fft_in', 'fft_info', 'fe_info', 'fe_derot', 'fe_interp', 'fe_decim'):
            f = getattr(L, 'ref_tap_' + n)
            f.argtypes = [C.c_void_p, C.c_longlong]
            f.restype = C.c_longlong
        _ref_chain = L
    return _ref_chain


class RefRx:
    """Reference demodulator-side stages for one transmission mode."""

    def __init__(self, fft='32K', carrier_ext=True, pp=7, gi='1/128', n_data=59, papr=0):
        self.L = ref_chain()
        out = np.zeros(13, np.int32)
        self.L.ref_rx_init(FFT_MODE[fft], 1 if carrier_ext else 0, pp - 1, GI[gi], n_data, papr, out)
        self.p = dict(zip(PARAM_NAMES, [int(x) for x in out]))

    def tables(self):
        """dict of the init-time tables the drop-in facade hands over to the GPU engine"""
        p, L = self.p, self.L
        k = p['k_total']
        t = {}
        nd = p['len_frame'] - p['l_fc'] - p['n_p2']
        t['data_map'] = np.zeros((nd, k), np.int32)
        t['data_ref'] = np.zeros((nd, k), np.float32)
        for i in range(nd):
            L.ref_rx_tables_data(i, t['data_map'][i], t['data_ref'][i])
        t['p2_map'] = np.zeros(k, np.int32)
        t['p2_ref'] = np.zeros(k, np.float32)
        L.ref_rx_tables_p2(t['p2_map'], t['p2_ref'])
        if p['l_fc']:
            t['fc_map'] = np.zeros(k, np.int32)
            t['fc_ref'] = np.zeros(k, np.float32)
            L.ref_rx_tables_fc(t['fc_map'], t['fc_ref'])
        for kind, name in enumerate(['p2', 'data', 'fc']):
            e, o = np.zeros(32768, np.int32), np.zeros(32768, np.int32)
            L.ref_rx_tables_h(kind, e, o)
            t['h_even_' + name], t['h_odd_' + name] = e, o
        a = [C.c_float() for _ in range(3)]
        L.ref_rx_amps(*[C.byref(x) for x in a])
        t['amp_p2'], t['amp_sp'], t['amp_cp'] = [x.value for x in a]
        return t

    def fft(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        out = np.empty_like(x)
        self.L.ref_fft(x.ctypes.data, out.ctypes.data)
        return out

    def data_symbol(self, idx_symbol, freq):
        freq = np.ascontiguousarray(freq, np.complex64)
        out = np.empty(self.p['c_data'], np.complex64)
        sro, ph = C.c_float(), C.c_float()
        self.L.ref_data_symbol(idx_symbol, freq.ctypes.data, out.ctypes.data, C.byref(sro), C.byref(ph))
        return out, sro.value, ph.value

    def fc_symbol(self, freq):
        freq = np.ascontiguousarray(freq, np.complex64)
        out = np.empty(self.p['n_fc'], np.complex64)
        sro, ph = C.c_float(), C.c_float()
        self.L.ref_fc_symbol(freq.ctypes.data, out.ctypes.data, C.byref(sro), C.byref(ph))
        return out, sro.value, ph.value

    def p2_symbol(self, freq):
        freq = np.ascontiguousarray(freq, np.complex64)
        out = np.empty(self.p['c_p2'], np.complex64)
        sro, ph = C.c_float(), C.c_float()
        crc = self.L.ref_p2_symbol(freq.ctypes.data, out.ctypes.data, C.byref(sro), C.byref(ph))
        return out, sro.value, ph.value, crc


def _tap(L, name, dtype):
    f = getattr(L, 'ref_tap_' + name)
    n = f(None, 0)
    a = np.empty(n, dtype)
    if n:
        f(a.ctypes.data, n)
    return a


class RefFec:
    """Reference FEC chain (time_deinterleaver -> llr_demapper -> ldpc_decoder -> bch_decoder -> bb_de_header)
    behind a RefRx (needs its dvbt2_parameters).  plps: list of dicts(id, cod, mod, rot, fec, blocks_max, ti_len, ti_type)."""

    def __init__(self, rx, plps, l1_post_size, need_plp=0):
        self.L = rx.L
        self.rx = rx
        d = np.array([[p['id'], p['cod'], p['mod'], p['rot'], p['fec'], p['blocks_max'], p['ti_len'], p['ti_type']]
                      for p in plps], np.int32)
        self.L.ref_fec_start(len(plps), np.ascontiguousarray(d.reshape(-1)), l1_post_size, need_plp)
        self.n_plp = len(plps)

    def chain(self, after_ti=True, after_demap=True, after_ldpc=True, after_bch=True):
        self.L.ref_fec_chain(int(after_ti), int(after_demap), int(after_ldpc), int(after_bch))

    def feed_p2(self, starts, num_blocks, cells):
        cells = np.ascontiguousarray(cells, np.complex64).copy()
        self.L.ref_fec_feed_p2(np.asarray(starts, np.int32), np.asarray(num_blocks, np.int32), len(cells), cells.ctypes.data)

    def feed(self, cells):
        cells = np.ascontiguousarray(cells, np.complex64).copy()
        self.L.ref_fec_feed(len(cells), cells.ctypes.data)

    def demap(self, cells, plp=0):
        cells = np.ascontiguousarray(cells, np.complex64).copy()
        self.L.ref_demap(len(cells), cells.ctypes.data, plp)

    def deheader(self, bits, plp=0):
        """one BBFRAME (uint8, one byte per bit) through the reference's bb_de_header::execute; the datagram is in taps()['ts']"""
        n = len(bits)
        # normal mode reads its per-packet CRC bytes without counting them against DFL (bb_de_header.cpp:286-296), i.e.
        # past the end of a full data field: give both implementations the same defined bytes there
        bits = np.concatenate([np.ascontiguousarray(bits, np.uint8), np.zeros(4096, np.uint8)])
        self.L.ref_bb_deheader(plp, n, bits.ctypes.data)

    def ldpc_batch(self, llr32, plp=0):
        """32 FECFRAMEs of int8 LLRs through the reference's ldpc_decoder::execute and the stages chained behind it"""
        llr32 = np.ascontiguousarray(llr32, np.int8)
        assert llr32.shape[0] == 32
        self.L.ref_ldpc_batch(plp, llr32.ctypes.data, llr32.shape[1])

    def permutation(self, plp=0):
        out = np.zeros(1 << 22, np.int32)
        n = self.L.ref_ti_permutation(plp, out, len(out))
        return out[:n].copy()

    def taps(self):
        L = self.L
        return {'ti_cells': _tap(L, 'ti_cells', np.complex64), 'ti_sizes': _tap(L, 'ti_sizes', np.int32),
                'llr': _tap(L, 'llr', np.int8), 'ldpc_bits': _tap(L, 'ldpc_bits', np.uint8),
                'bb_bits': _tap(L, 'bb_bits', np.uint8), 'bb_len': _tap(L, 'bb_len', np.int32),
                'snr': _tap(L, 'snr', np.float32), 'ts': _tap(L, 'ts', np.uint8),
                'ts_datagrams': _tap(L, 'ts_datagrams', np.int32)}

    def clear(self):
        self.L.ref_tap_clear()


# ---- port oracle: FEC front half (oracle/port/fec_port.c) ----
def port_cell_permutation(n_fec, cells_per_fec):
    out = np.zeros(n_fec * cells_per_fec, np.int32)
    port().port_cell_permutation(n_fec, cells_per_fec, out.ctypes.data)
    return out


def port_ti_blocks(cells, n_fec_per_block, cells_per_fec, perm, state=None):
    """cells complex64 stream of consecutive TI blocks -> de-interleaved blocks.  state = [end_cell, q_first]
    carried like the reference's members (zero for a fresh receiver)."""
    cells = np.ascontiguousarray(cells, np.complex64)
    out = np.zeros_like(cells)
    end_cell, q_first = C.c_int(0 if state is None else state[0]), C.c_float(0.0 if state is None else state[1])
    perm = np.ascontiguousarray(perm, np.int32)
    off = 0
    for nf in n_fec_per_block:
        n = nf * cells_per_fec
        src, dst = cells[off:off + n], out[off:off + n]
        port().port_ti_block(src.ctypes.data, nf, cells_per_fec, perm.ctypes.data, dst.ctypes.data,
                             C.byref(end_cell), C.byref(q_first))
        off += n
    if state is not None:
        state[0], state[1] = end_cell.value, q_first.value
    return out


def port_demap_address(fec_normal, mod, code_rate):
    out = np.zeros(64800 if fec_normal else 16200, np.int32)
    port().port_demap_address_for(int(fec_normal), mod, code_rate, out.ctypes.data)
    return out


def port_demap(cells, mod, rotation, fec_normal, code_rate, precision_in=0.0, saturate=False):
    """one TI block; returns (llr int8[n_fec][N], snr, precision, derotated cells).  saturate: the PRODUCT's
    T2B200_OPT_DEMAP_SATURATE cast instead of the reference's wrapping one"""
    port().port_set_demap_saturate(int(saturate))
    cells = np.ascontiguousarray(cells, np.complex64).copy()
    nb = 64800 if fec_normal else 16200
    cpf = nb // (2 * (mod + 1))
    llr = np.zeros((len(cells) // cpf, nb), np.int8)
    snr = C.c_float()
    p = port().port_demap(cells.ctypes.data, len(cells), mod, int(rotation), int(fec_normal), code_rate,
                          llr.ctypes.data, C.byref(snr), float(precision_in))
    return llr, snr.value, p, cells


# ---- port oracle: FFT + equaliser (oracle/port/eq_port.c) ----
def port_equalize(kind, freq, l_nulls, k_total, cmap, refer, h, n_out, amp_main, amp_cp=0.0):
    """one symbol; kind 0 P2 / 1 data / 2 FC -> (cells complex64[n_out], sro, phase)"""
    freq = np.ascontiguousarray(freq, np.complex64)
    cmap = np.ascontiguousarray(cmap, np.int32)
    refer = np.ascontiguousarray(refer, np.float32)
    h = np.ascontiguousarray(h, np.int32)
    out = np.zeros(n_out, np.complex64)
    sro, ph = C.c_float(), C.c_float()
    port().port_equalize(kind, freq.ctypes.data, l_nulls, k_total, cmap.ctypes.data, refer.ctypes.data, h.ctypes.data,
                         float(amp_main), float(amp_cp), out.ctypes.data, C.byref(sro), C.byref(ph))
    return out, sro.value, ph.value


def port_fft(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    port().port_fft_shift(x.ctypes.data, len(x), out.ctypes.data)
    return out


# ---- port oracle: BBFRAME -> TS re-packetiser (oracle/port/ts_port.c) ----
class PortTs:
    """stateful restatement of bb_de_header::execute: feed(bits) -> datagram bytes (None when the frame is dropped)"""

    def __init__(self):
        L = port()
        L.port_ts_state_size.restype = C.c_int
        L.port_ts_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.port_ts_frame.restype = C.c_int
        L.port_ts_reset.argtypes = [C.c_void_p]
        self.L = L
        self.state = np.zeros(L.port_ts_state_size(), np.uint8)
        L.port_ts_reset(self.state.ctypes.data)

    def feed(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        padded = np.concatenate([bits, np.zeros(70000, np.uint8)])     # normal mode reads its CRC bytes beyond DFL; a frame entered with
        # a packet index beyond 188 reads far past the frame (undefined in the reference): zeros here, as on the GPU
        out = np.zeros(len(bits) // 8 + 2 * 188 + 64, np.uint8)
        n = self.L.port_ts_frame(self.state.ctypes.data, padded.ctypes.data, len(bits), out.ctypes.data)
        return None if n < 0 else out[:n].copy()


class PortFrontend:
    """oracle/port/frontend_port.c: the receiver front-end of one stream, chunk by chunk (dvbt2_demodulator.cpp:178-221).
    The state is exposed as a numpy record so that a test can start from a state the reference reported."""
    STATE = np.dtype([('dc_re', 'f4'), ('dc_im', 'f4'), ('frequency_nco', 'f4'), ('x1', 'f4'), ('delay', 'f4', (3, 2)),
                      ('hist', 'f4', (63, 2)), ('parity', 'i4')])

    def __init__(self):
        self.L = port()
        assert self.L.port_fe_state_size() == self.STATE.itemsize
        self.state = np.zeros(1, self.STATE)
        self.L.port_fe_reset(self.state.ctypes.data)
        self.theta = np.zeros(3, np.float32)

    def chunk(self, i16, q16, short_to_float, c1, c2, frequency_est_filtered, phase_nco, resample, stride=1):
        """-> (decimator output complex64, resampler output complex64, derotated samples complex64)"""
        i16 = np.ascontiguousarray(i16, np.int16)
        q16 = np.ascontiguousarray(q16, np.int16)
        n = len(i16) // stride
        derot = np.empty(n, np.complex64)
        interp = np.empty(int(n / max(float(np.float32(resample)), 0.2)) + 8, np.complex64)
        out = np.empty(len(interp) // 2 + 2, np.complex64)
        n_interp = C.c_int()
        k = self.L.port_fe_chunk(self.state.ctypes.data, i16.ctypes.data, q16.ctypes.data, stride, n, short_to_float, c1, c2,
                                 frequency_est_filtered, phase_nco, resample, derot.ctypes.data, interp.ctypes.data,
                                 C.byref(n_interp), out.ctypes.data, self.theta.ctypes.data)
        return out[:k].copy(), interp[:n_interp.value].copy(), derot


class PortP1:
    """oracle/port/frontend_port.c: p1_symbol's sliding correlator (p1_symbol.cpp:75-178), state carried between calls"""

    def __init__(self):
        self.L = port()
        self.state = np.zeros(self.L.port_p1_state_size(), np.uint8)
        self.L.port_p1_reset(self.state.ctypes.data)

    def correlate(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        corr = np.empty(len(x), np.float32)
        out = np.empty(len(x), np.complex64)
        self.L.port_p1_correlate(self.state.ctypes.data, x.ctypes.data, len(x), corr.ctypes.data, out.ctypes.data)
        return corr, out


def ref_p1_trace(x):
    """the reference's own correlator, sample by sample (oracle/ref_chain.cc::ref_p1_trace) -> correlation float32[n]"""
    x = np.ascontiguousarray(x, np.complex64)
    corr = np.empty(len(x), np.float32)
    ref_chain().ref_p1_trace(x.ctypes.data, len(x), corr.ctypes.data)
    return corr


def port_cp_correlate(sym, fft_size, guard):
    sym = np.ascontiguousarray(sym, np.complex64)
    return float(port().port_fe_cp_correlate(sym.ctypes.data, fft_size, guard))


class RefDemod:
    """The reference's whole receiver, dvbt2_demodulator::execute (dvbt2_demodulator.cpp:145-254) down to the TS sink, fed
    with int16 I/Q in front-end sized chunks.  One instance per process (the stages keep static state).  Besides the
    taps of RefFec it records the FFT input window of every OFDM symbol (`in_fft`, dvbt2_demodulator.cpp:332)."""

    def __init__(self, sample_rate=64e6 / 7, need_plp=0, tap_fft=True, tap_frontend=(0, 0)):
        self.L = ref_chain()
        self.L.ref_demod_new(float(sample_rate), need_plp)
        self.L.ref_tap_frontend_arm(int(tap_frontend[0]), int(tap_frontend[1]))
        self.L.ref_tap_fft_arm(1 if tap_fft else 0)
        self.status = []

    def feed(self, i16, q16, chunk=1 << 16):
        i16 = np.ascontiguousarray(i16, np.int16)
        q16 = np.ascontiguousarray(q16, np.int16)
        for a in range(0, len(i16), chunk):
            n = min(chunk, len(i16) - a)
            self.status.append(self.L.ref_demod_feed(n, i16[a:a + n].ctypes.data, q16[a:a + n].ctypes.data))

    FE_INFO = ['len_in', 'len_interp', 'resample', 'phase_nco', 'frequency_est_filtered', 'c1', 'c2', 'frequency_nco_after',
               'dc_re_after', 'dc_im_after', 'len_out', 'short_to_float', 'x1_after', 'n_interp_after']

    def frontend_taps(self):
        """Per chunk of dvbt2_demodulator::execute (oracle/ref_chain.cc, oracle_tap_farrow): the loop parameters and state
        (FE_INFO) of every chunk; for the chunks of the armed window (first, count) the derotated samples (resampler input),
        the resampler output and the decimator output, back to back."""
        info = _tap(self.L, 'fe_info', np.float64).reshape(-1, 16)[:, :14]
        info[:, 10] = np.floor(info[:, 10])
        return {'info': info, 'derot': _tap(self.L, 'fe_derot', np.complex64), 'interp': _tap(self.L, 'fe_interp', np.complex64),
                'decim': _tap(self.L, 'fe_decim', np.complex64)}

    def params(self):
        out = np.zeros(16, np.int32)
        self.L.ref_demod_params(out)
        return dict(zip(PARAM_NAMES + ['pilot_pattern', 'l1_post_size', 'num_plp'], [int(x) for x in out]))

    def taps(self):
        L = self.L
        t = {'ti_sizes': _tap(L, 'ti_sizes', np.int32), 'bb_bits': _tap(L, 'bb_bits', np.uint8),
             'bb_len': _tap(L, 'bb_len', np.int32), 'snr': _tap(L, 'snr', np.float32), 'ts': _tap(L, 'ts', np.uint8),
             'ts_datagrams': _tap(L, 'ts_datagrams', np.int32), 'fft_info': _tap(L, 'fft_info', np.int32).reshape(-1, 3)}
        x = _tap(L, 'fft_in', np.complex64)
        t['fft_in'] = x.reshape(len(t['fft_info']), -1) if len(t['fft_info']) else x
        return t


# ---- the drop-in proof: reference receiver front half + GPU stage classes (oracle/dropin_glue.cc) ----
class DropinDemod:
    """The reference's UNMODIFIED dvbt2_demodulator / p1_symbol / p2_symbol / bb_de_header compiled against the GPU stage
    classes of sdr_receiver_dvb_t2_b200/host/dropin (oracle/_ref/libdropin_chain.so): int16 I/Q in, TS datagrams out.
    Needs a GPU (the stage classes have no CPU fallback).  One instance per process."""

    def __init__(self, sample_rate=64e6 / 7, need_plp=0, gpu_frontend=False):
        """gpu_frontend: the per-sample loop, resampler and decimator of dvbt2_demodulator::execute run on the GPU as well
        (t2b200_frontend_execute, one call per chunk; oracle/dropin_glue.cc::dropin_demod_feed_gpu_frontend)"""
        L = C.CDLL(os.path.join(_HERE, '_ref', 'libdropin_chain.so'))
        L.dropin_demod_new.argtypes = [C.c_float, C.c_int]
        L.dropin_demod_feed.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.dropin_demod_feed_gpu_frontend.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        self._feed = L.dropin_demod_feed_gpu_frontend if gpu_frontend else L.dropin_demod_feed
        L.dropin_launches.restype = C.c_longlong
        for n in ('ts', 'ts_datagrams', 'bb_bits', 'bb_len'):
            f = getattr(L, 'dropin_tap_' + n)
            f.argtypes = [C.c_void_p, C.c_longlong]
            f.restype = C.c_longlong
        self.L = L
        L.dropin_demod_new(float(sample_rate), need_plp)
        self.status = []

    def feed(self, i16, q16, chunk=1 << 16):
        i16 = np.ascontiguousarray(i16, np.int16)
        q16 = np.ascontiguousarray(q16, np.int16)
        for a in range(0, len(i16), chunk):
            n = min(chunk, len(i16) - a)
            self.status.append(self._feed(n, i16[a:a + n].ctypes.data, q16[a:a + n].ctypes.data))
            assert self.status[-1] >= 0

    def taps(self):
        def tap(name, dtype):
            f = getattr(self.L, 'dropin_tap_' + name)
            n = f(None, 0)
            a = np.empty(n, dtype)
            if n:
                f(a.ctypes.data, n)
            return a
        t = {'ts': tap('ts', np.uint8), 'ts_datagrams': tap('ts_datagrams', np.int32), 'bb_bits': tap('bb_bits', np.uint8),
             'bb_len': tap('bb_len', np.int32), 'launches': int(self.L.dropin_launches())}
        n = len(t['bb_len'])
        t['bb_bits'] = t['bb_bits'].reshape(n, -1) if n else t['bb_bits']
        return t
