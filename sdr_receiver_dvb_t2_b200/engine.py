"""ctypes binding of include/t2b200.h.

Arrays may be numpy arrays (host memory) or torch CUDA tensors (device memory, passed by
data_ptr()); the C library detects which.  No computation happens here.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NOMEM = range(5)
LDPC_GROUP32, LDPC_BCH_DESCRAMBLE, LDPC_PACK_BITS, LDPC_WANT_POST = 1, 2, 4, 8
C1_2, C3_5, C2_3, C3_4, C4_5, C5_6 = range(6)
MOD_QPSK, MOD_16QAM, MOD_64QAM, MOD_256QAM = range(4)
FEC_SHORT, FEC_NORMAL = 0, 1
OPT_DEMAP_SATURATE = 1
OPT_LDPC_PLAIN_LAUNCH = 2
OPT_BCH_CORRECT = 3
OPT_STAGE_TIMING = 4

# every symbol include/t2b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    't2b200_create', 't2b200_destroy', 't2b200_set_stream', 't2b200_sync', 't2b200_last_error',
    't2b200_version', 't2b200_launch_count', 't2b200_set_option',
    't2b200_ldpc_code_id', 't2b200_ldpc_n', 't2b200_ldpc_k', 't2b200_ldpc_k_bch',
    't2b200_ldpc_decode', 't2b200_bch_descramble',
    't2b200_cell_permutation', 't2b200_demap_address_table', 't2b200_freq_deinterleaver_table', 't2b200_ti_configure', 't2b200_ti_deinterleave',
    't2b200_demap', 't2b200_eq_configure', 't2b200_equalize', 't2b200_fft',
    't2b200_ts_reset', 't2b200_ts_packetize', 't2b200_frames_configure', 't2b200_frames_decode',
    't2b200_mode_init', 't2b200_pilot_tables', 't2b200_eq_configure_mode', 't2b200_frames_decode_i16',
    't2b200_comm_unique_id', 't2b200_comm_init', 't2b200_comm_destroy', 't2b200_ldpc_decode_sharded',
    't2b200_bch_t', 't2b200_bch_decode', 't2b200_frames_stage_ms',
    't2b200_frontend_configure', 't2b200_frontend_reset', 't2b200_frontend_execute', 't2b200_frontend_get_state',
    't2b200_frontend_set_state', 't2b200_cp_correlate', 't2b200_p1_correlate',
]


class FrameCfg(C.Structure):
    """t2b200_frame_cfg"""
    _fields_ = [(n, C.c_int) for n in ('fft_size', 'len_frame', 'n_p2', 'l_fc', 'c_p2', 'c_data', 'n_fc', 'first_cell',
                                       'plp', 'mod', 'rotation', 'fec_type', 'code_rate', 'n_blocks', 'ti_len')]


class Mode(C.Structure):
    """t2b200_mode"""
    _fields_ = [(n, C.c_int) for n in ('fft_mode', 'carrier_mode', 'pilot_pattern', 'guard_interval_mode', 'papr_mode',
                                       'fft_size', 'k_total', 'k_ext', 'k_offset', 'l_nulls', 'guard_interval_size',
                                       'n_p2', 'c_p2', 'c_data', 'n_fc', 'c_fc', 'l_fc', 'n_data', 'len_frame', 'dx', 'dy')] + \
               [(n, C.c_float) for n in ('amp_p2', 'amp_sp', 'amp_cp')]


# t2b200_fe_state / t2b200_fe_chunk / t2b200_fe_result as numpy record types
FE_STATE = np.dtype([('dc_re', 'f4'), ('dc_im', 'f4'), ('frequency_nco', 'f4'), ('x1', 'f4'), ('delay', 'f4', (3, 2)),
                     ('hist', 'f4', (63, 2)), ('parity', 'i4'), ('pad', 'i4')])
FE_CHUNK = np.dtype([('len_in', 'i4'), ('short_to_float', 'f4'), ('c1', 'f4'), ('c2', 'f4'), ('frequency_est_filtered', 'f4'),
                     ('phase_nco', 'f4'), ('resample', 'f4')])
FE_RESULT = np.dtype([('len_out', 'i4'), ('len_interp', 'i4'), ('theta', 'f4', (3,))])

FFT_MODE = {'16K': 4, '32K': 5}
GI_MODE = {'1/32': 0, '1/16': 1, '1/8': 2, '1/4': 3, '1/128': 4, '19/128': 5, '19/256': 6}


class T2Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, 'libt2b200.so')


def lib():
    """Load (building first if a source is newer) libt2b200.so.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        _build.build()
    except Exception as e:  # no nvcc on the GPU box: the prebuilt library must be there
        if not os.path.exists(lib_path()):
            raise T2Error('libt2b200.so is missing and cannot be built: %s' % e)
    L = C.CDLL(lib_path())
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint
    L.t2b200_create.argtypes = [i32, C.POINTER(vp)]
    L.t2b200_destroy.argtypes = [vp]
    L.t2b200_destroy.restype = None
    L.t2b200_set_stream.argtypes = [vp, vp]
    L.t2b200_sync.argtypes = [vp]
    L.t2b200_last_error.argtypes = [vp]
    L.t2b200_last_error.restype = C.c_char_p
    L.t2b200_version.restype = C.c_char_p
    L.t2b200_launch_count.argtypes = [vp]
    L.t2b200_set_option.argtypes = [vp, i32, i32]
    L.t2b200_launch_count.restype = C.c_longlong
    L.t2b200_ldpc_decode.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, i32, u32]
    L.t2b200_bch_descramble.argtypes = [vp, i32, vp, i32, vp]
    L.t2b200_bch_decode.argtypes = [vp, i32, vp, i32, vp]
    L.t2b200_bch_t.argtypes = [i32]
    L.t2b200_cell_permutation.argtypes = [i32, i32, vp]
    L.t2b200_demap_address_table.argtypes = [i32, i32, i32, vp]
    L.t2b200_freq_deinterleaver_table.argtypes = [i32, i32, vp, vp]
    L.t2b200_ti_configure.argtypes = [vp, i32, i32, i32, i32, vp]
    L.t2b200_ti_deinterleave.argtypes = [vp, i32, vp, i32, vp, vp]
    L.t2b200_demap.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.t2b200_eq_configure.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, C.c_float, C.c_float]
    L.t2b200_equalize.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.t2b200_fft.argtypes = [vp, i32, vp, i32, vp]
    L.t2b200_ts_reset.argtypes = [vp, i32]
    L.t2b200_ts_packetize.argtypes = [vp, i32, vp, i32, i32, vp, C.c_size_t, vp, vp, C.POINTER(C.c_longlong)]
    L.t2b200_frames_configure.argtypes = [vp, C.POINTER(FrameCfg)]
    L.t2b200_frames_decode.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, u32]
    L.t2b200_frames_decode_i16.argtypes = [vp, vp, C.c_float, i32, vp, vp, vp, vp, vp, i32, u32]
    L.t2b200_frames_stage_ms.argtypes = [vp, vp]
    L.t2b200_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.t2b200_comm_init.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.t2b200_comm_destroy.argtypes = [vp]
    L.t2b200_ldpc_decode_sharded.argtypes = [vp, i32, i32, vp, i32, vp, i32, u32]
    L.t2b200_frontend_configure.argtypes = [vp, i32, i32]
    L.t2b200_frontend_reset.argtypes = [vp, i32]
    L.t2b200_frontend_execute.argtypes = [vp, vp, vp, C.c_longlong, i32, vp, vp, C.c_longlong, vp]
    L.t2b200_frontend_get_state.argtypes = [vp, i32, vp]
    L.t2b200_frontend_set_state.argtypes = [vp, i32, vp]
    L.t2b200_cp_correlate.argtypes = [vp, vp, i32, C.c_longlong, i32, i32, vp]
    L.t2b200_p1_correlate.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.t2b200_mode_init.argtypes = [i32] * 6 + [C.POINTER(Mode)]
    L.t2b200_pilot_tables.argtypes = [C.POINTER(Mode), i32, vp, vp]
    L.t2b200_eq_configure_mode.argtypes = [vp, C.POINTER(Mode)]
    _lib = L
    return L


def _ptr(a):
    """address of a numpy array / torch tensor / None"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags['C_CONTIGUOUS']
        return a.ctypes.data
    if hasattr(a, 'data_ptr'):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def _is_torch(a):
    return hasattr(a, 'data_ptr') and not isinstance(a, np.ndarray)


def _like(ref, shape, dtype_np):
    """allocate an output next to `ref`: numpy -> numpy, torch cuda tensor -> torch cuda tensor"""
    if _is_torch(ref):
        import torch
        td = {np.uint8: torch.uint8, np.int8: torch.int8, np.int32: torch.int32,
              np.float32: torch.float32, np.complex64: torch.complex64, np.int16: torch.int16}[dtype_np]
        return torch.empty(shape, dtype=td, device=ref.device)
    return np.empty(shape, dtype_np)


class Engine:
    """One t2b200 context = one GPU + one stream."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.t2b200_create(device, C.byref(h))
        if rc != OK:
            raise T2Error('t2b200_create(device=%d) failed (rc=%d): no usable CUDA device; '
                          'there is no CPU fallback' % (device, rc))
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, 'h', None):
            self.L.t2b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != OK:
            raise T2Error('t2b200 rc=%d: %s' % (rc, self.L.t2b200_last_error(self.h).decode()))

    def sync(self):
        self._chk(self.L.t2b200_sync(self.h))

    def set_stream(self, stream):
        """stream: a cudaStream_t handle as int (torch: stream.cuda_stream); 0 = the legacy default
        stream (what torch's default stream is), None = back to the context's own stream"""
        if stream is None:
            h = 0
        elif stream == 0:
            h = 1          # cudaStreamLegacy
        else:
            h = stream
        self._chk(self.L.t2b200_set_stream(self.h, C.c_void_p(h)))

    def set_option(self, option, value):
        self._chk(self.L.t2b200_set_option(self.h, option, int(value)))

    def frames_stage_ms(self):
        """device ms of every stage of the last frames_decode call made with OPT_STAGE_TIMING on (t2b200_frames_stage_ms)"""
        ms = (C.c_float * 6)()
        self._chk(self.L.t2b200_frames_stage_ms(self.h, ms))
        return dict(zip(('fft', 'equalize', 'ti_deinterleave', 'demap', 'ldpc_bch', 'call'), [float(x) for x in ms]))

    # ---- N2: receiver front-end ----
    def frontend_configure(self, n_streams, max_chunk_in):
        self._chk(self.L.t2b200_frontend_configure(self.h, n_streams, max_chunk_in))
        self._fe_streams = n_streams

    def frontend_reset(self, stream=-1):
        self._chk(self.L.t2b200_frontend_reset(self.h, stream))

    def frontend_state(self, stream):
        st = np.zeros(1, FE_STATE)
        self._chk(self.L.t2b200_frontend_get_state(self.h, stream, st.ctypes.data))
        return st[0]

    def frontend_set_state(self, stream, state):
        st = np.zeros(1, FE_STATE)
        st[0] = state
        self._chk(self.L.t2b200_frontend_set_state(self.h, stream, st.ctypes.data))

    def frontend_execute(self, i_in, q_in, chunks, out=None, stream_stride=None, sample_step=1):
        """i_in, q_in: int16 [n_streams][...] (numpy or torch cuda); chunks: FE_CHUNK[n_streams].
        -> (out complex64 [n_streams][out_stride], results FE_RESULT[n_streams])"""
        n = self._fe_streams
        chunks = np.ascontiguousarray(chunks, FE_CHUNK)
        assert len(chunks) == n
        if stream_stride is None:
            stream_stride = int(i_in.shape[1]) if len(i_in.shape) > 1 else 0
        if out is None:
            worst = int((int(chunks['len_in'].max()) + 1) / float(chunks['resample'].min()) + 2) // 2 + 2
            out = _like(i_in, (n, worst), np.complex64)
        res = np.zeros(n, FE_RESULT)
        self._chk(self.L.t2b200_frontend_execute(self.h, _ptr(i_in), _ptr(q_in), stream_stride, sample_step, chunks.ctypes.data,
                                                 _ptr(out), int(out.shape[1]), res.ctypes.data))
        return out, res

    def cp_correlate(self, symbols, fft_size, guard):
        """symbols: complex64 [n][guard + fft_size] -> frequency_est float32[n] (dvbt2_demodulator.cpp:321-330)"""
        est = _like(symbols, (symbols.shape[0],), np.float32)
        self._chk(self.L.t2b200_cp_correlate(self.h, _ptr(symbols), int(symbols.shape[0]), int(symbols.shape[1]), fft_size, guard,
                                             _ptr(est)))
        return est

    def p1_correlate(self, samples, history=None, fq_index=0):
        """samples complex64[n], history complex64[2046] or None -> (correlation float32[n], out complex64[n]) (p1_symbol.cpp:75-178)"""
        n = int(samples.shape[0])
        corr = _like(samples, (n,), np.float32)
        out = _like(samples, (n,), np.complex64)
        self._chk(self.L.t2b200_p1_correlate(self.h, _ptr(samples), n, _ptr(history), fq_index, _ptr(corr), _ptr(out)))
        return corr, out

    @property
    def launches(self):
        return self.L.t2b200_launch_count(self.h)

    # ---- LDPC (+ BCH strip / descramble) ----
    def ldpc_code_id(self, fec_type, code_rate):
        return self.L.t2b200_ldpc_code_id(fec_type, code_rate)

    def ldpc_geometry(self, code):
        return self.L.t2b200_ldpc_n(code), self.L.t2b200_ldpc_k(code), self.L.t2b200_ldpc_k_bch(code)

    def ldpc_decode(self, code, llr, flags=LDPC_GROUP32, max_trials=25, want_status=True, out=None):
        """llr int8[n][N] -> dict(bits, trials_left, iterations[, post])"""
        N, K, KB = self.ldpc_geometry(code)
        n = llr.shape[0]
        assert llr.shape[1] == N
        k_out = KB if flags & LDPC_BCH_DESCRAMBLE else K
        row = k_out // 8 if flags & LDPC_PACK_BITS else k_out
        bits = out if out is not None else _like(llr, (n, row), np.uint8)
        tl = _like(llr, (n,), np.int32) if want_status else None
        it = _like(llr, (n,), np.int32) if want_status else None
        post = _like(llr, (n, N), np.int8) if flags & LDPC_WANT_POST else None
        self._chk(self.L.t2b200_ldpc_decode(self.h, code, _ptr(llr), n, _ptr(bits), _ptr(tl), _ptr(it), _ptr(post),
                                            max_trials, flags))
        r = {'bits': bits, 'trials_left': tl, 'iterations': it}
        if post is not None:
            r['post'] = post
        return r

    # ---- multi-GPU: the FEC stage sharded by codeword (t2b200_comm_*, t2b200_ldpc_decode_sharded) ----
    def comm_init(self, rank, nranks, unique_id):
        """unique_id: the 128 bytes rank 0 got from comm_unique_id(), the same on every rank"""
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._chk(self.L.t2b200_comm_init(self.h, rank, nranks, buf, COMM_ID_BYTES))

    def comm_destroy(self):
        self._chk(self.L.t2b200_comm_destroy(self.h))

    def ldpc_decode_sharded(self, code, root, llr, n_cw, out=None, flags=LDPC_GROUP32 | LDPC_BCH_DESCRAMBLE, max_trials=25):
        """collective: llr (torch cuda int8[n_cw][N]) and out live on `root`, the other ranks pass None"""
        self._chk(self.L.t2b200_ldpc_decode_sharded(self.h, code, root, _ptr(llr), n_cw, _ptr(out), max_trials, flags))
        return out

    def bch_decode(self, code, bits):
        """opt-in N3: bits uint8[n][K_ldpc] (byte per bit) corrected IN PLACE; returns int32[n] errors corrected (-1: > t)"""
        n = bits.shape[0]
        cor = _like(bits, (n,), np.int32)
        self._chk(self.L.t2b200_bch_decode(self.h, code, _ptr(bits), n, _ptr(cor)))
        return cor

    def bch_descramble(self, code, bits):
        N, K, KB = self.ldpc_geometry(code)
        n = bits.shape[0]
        assert bits.shape[1] == K
        out = _like(bits, (n, KB), np.uint8)
        self._chk(self.L.t2b200_bch_descramble(self.h, code, _ptr(bits), n, _ptr(out)))
        return out

    # ---- K3 / K4 ----
    def ti_configure(self, plp, fec_type, mod, n_fec_blocks_max, permutation=None):
        if permutation is not None:
            permutation = np.ascontiguousarray(permutation, np.int32)
        self._chk(self.L.t2b200_ti_configure(self.h, plp, fec_type, mod, n_fec_blocks_max, _ptr(permutation)))

    def ti_deinterleave(self, plp, cells, n_fec_per_block):
        """cells complex64[total cells] (numpy or torch cuda) -> de-interleaved TI blocks, same type"""
        nf = np.ascontiguousarray(n_fec_per_block, np.int32)
        out = _like(cells, tuple(cells.shape), np.complex64)
        self._chk(self.L.t2b200_ti_deinterleave(self.h, plp, _ptr(cells), len(nf), _ptr(nf), _ptr(out)))
        return out

    def demap(self, ti_cells, n_fec_per_block, mod, rotation, fec_type, code_rate, precision_in=None):
        """ti_cells is derotated IN PLACE when rotation != 0 (as the reference does).
        -> dict(llr int8[n_fec][N], snr, precision)"""
        nf = np.ascontiguousarray(n_fec_per_block, np.int32)
        n_fec, nbits = int(nf.sum()), (64800 if fec_type else 16200)
        llr = _like(ti_cells, (n_fec, nbits), np.int8)
        snr = _like(ti_cells, (len(nf),), np.float32)          # device buffers in, device buffers out: the call stays asynchronous
        prec = _like(ti_cells, (len(nf),), np.float32)
        if precision_in is not None:
            precision_in = np.ascontiguousarray(precision_in, np.float32)
        self._chk(self.L.t2b200_demap(self.h, _ptr(ti_cells), len(nf), _ptr(nf), mod, rotation, fec_type, code_rate,
                                      _ptr(llr), _ptr(snr), _ptr(prec), _ptr(precision_in)))
        return {'llr': llr, 'snr': snr, 'precision': prec}

    # ---- N1: BBFRAME -> TS ----
    def ts_reset(self, plp=0):
        self._chk(self.L.t2b200_ts_reset(self.h, plp))

    def ts_packetize(self, bbframes, plp=0):
        """bbframes uint8[n][k_bch] (one byte per bit; numpy / torch cuda) -> (ts bytes, datagram_len[n], status[n])"""
        n, k_bch = bbframes.shape
        cap = n * (k_bch // 8 + 2 * 188)
        ts = _like(bbframes, (cap,), np.uint8)
        dl, st = np.zeros(n, np.int32), np.zeros(n, np.int32)
        total = C.c_longlong(0)
        self._chk(self.L.t2b200_ts_packetize(self.h, plp, _ptr(bbframes), n, k_bch, _ptr(ts), cap, _ptr(dl), _ptr(st),
                                             C.byref(total)))
        return ts[:total.value], dl, st

    # ---- whole frames in one call ----
    def frames_configure(self, **kw):
        cfg = FrameCfg(**kw)
        self._frame_cfg = cfg
        self._chk(self.L.t2b200_frames_configure(self.h, C.byref(cfg)))

    def frames_decode(self, iq, flags=LDPC_GROUP32 | LDPC_BCH_DESCRAMBLE, max_trials=25, want_status=True, out=None, scale=None):
        """iq complex64[F][len_frame][fft_size] (numpy / pinned / torch cuda) -> dict(bits, trials_left, sro, phase, snr);
        outputs live where iq lives (torch cuda in -> torch cuda out, nothing waits for the GPU).
        scale given: iq is int16[F][len_frame][fft_size][2] (I, Q pairs), sample = (I + jQ) * scale (t2b200_frames_decode_i16)"""
        c = self._frame_cfg
        F = iq.shape[0]
        n_cw = F * c.n_blocks
        code = self.ldpc_code_id(c.fec_type, c.code_rate)
        N, K, KB = self.ldpc_geometry(code)
        k_out = KB if flags & LDPC_BCH_DESCRAMBLE else K
        row = k_out // 8 if flags & LDPC_PACK_BITS else k_out
        bits = out if out is not None else _like(iq, (n_cw, row), np.uint8)
        r = {'bits': bits, 'trials_left': None, 'sro': None, 'phase': None, 'snr': None}
        if want_status:
            r['trials_left'] = _like(iq, (n_cw,), np.int32)
            r['sro'], r['phase'] = _like(iq, (F, c.len_frame), np.float32), _like(iq, (F, c.len_frame), np.float32)
            r['snr'] = _like(iq, (F * c.ti_len,), np.float32)
        if scale is not None:
            self._chk(self.L.t2b200_frames_decode_i16(self.h, _ptr(iq), float(scale), F, _ptr(bits), _ptr(r['trials_left']),
                                                      _ptr(r['sro']), _ptr(r['phase']), _ptr(r['snr']), max_trials, flags))
        else:
            self._chk(self.L.t2b200_frames_decode(self.h, _ptr(iq), F, _ptr(bits), _ptr(r['trials_left']), _ptr(r['sro']),
                                                  _ptr(r['phase']), _ptr(r['snr']), max_trials, flags))
        return r

    # ---- K1 ----
    def fft(self, x, out=None):
        """x complex64[batch][n] (numpy / torch cuda) -> shifted unnormalised forward DFT, same type"""
        batch, n = x.shape
        o = out if out is not None else _like(x, (batch, n), np.complex64)
        self._chk(self.L.t2b200_fft(self.h, n, _ptr(x), batch, _ptr(o)))
        return o

    # ---- K2 ----
    def eq_configure(self, kind, first_symbol, fft_size, k_total, l_nulls, n_out, carrier_map, pilot_refer,
                     h_even, h_odd, amp_main, amp_cp=0.0):
        cm = np.ascontiguousarray(carrier_map, np.int32).reshape(-1, k_total)
        pr = np.ascontiguousarray(pilot_refer, np.float32).reshape(-1, k_total)
        he, ho = np.ascontiguousarray(h_even, np.int32), np.ascontiguousarray(h_odd, np.int32)
        self._eq_nout = getattr(self, '_eq_nout', {})
        self._eq_nout[kind] = (n_out, fft_size)
        self._chk(self.L.t2b200_eq_configure(self.h, kind, cm.shape[0], first_symbol, fft_size, k_total, l_nulls, n_out,
                                             _ptr(cm), _ptr(pr), _ptr(he), _ptr(ho), amp_main, amp_cp))

    def equalize(self, kind, idx_symbol, freq, out=None):
        """freq complex64[n][fft_size] (numpy / torch cuda) -> (cells complex64[n][n_out], sro[n], phase[n])"""
        n_out, fft_size = self._eq_nout[kind]
        idx = np.ascontiguousarray(idx_symbol, np.int32)
        n = len(idx)
        cells = out if out is not None else _like(freq, (n, n_out), np.complex64)
        sro, ph = _like(freq, (n,), np.float32), _like(freq, (n,), np.float32)
        self._chk(self.L.t2b200_equalize(self.h, kind, n, _ptr(idx), _ptr(freq), _ptr(cells), _ptr(sro), _ptr(ph)))
        return cells, sro, ph


def mode_init(fft='32K', carrier_ext=True, pp=7, gi='1/128', n_data=59, papr=0):
    """t2b200_mode_init -> Mode (dvbt2_parameters of one SISO transmission mode, built natively)"""
    m = Mode()
    rc = lib().t2b200_mode_init(FFT_MODE[fft], 1 if carrier_ext else 0, pp - 1, GI_MODE[gi], n_data, papr, C.byref(m))
    if rc != OK:
        raise T2Error('t2b200_mode_init rc=%d (combination not defined)' % rc)
    return m


def mode_tables(m):
    """The init-time tables of a mode, built natively (t2b200_pilot_tables + t2b200_freq_deinterleaver_table), in the
    layout oracle RefRx.tables() / tools.make_golden_tables.load() use: dict with 'p' (parameters), *_map, *_ref,
    h_even_* / h_odd_*, amp_*."""
    L = lib()
    k = m.k_total
    nd = m.len_frame - m.l_fc - m.n_p2
    t = {'p': {n: int(getattr(m, n)) for n in ('fft_size', 'k_total', 'l_nulls', 'c_p2', 'c_data', 'n_fc', 'c_fc', 'n_data',
                                               'len_frame', 'l_fc', 'n_p2', 'guard_interval_size', 'k_ext')}}
    t['amp_p2'], t['amp_sp'], t['amp_cp'] = float(m.amp_p2), float(m.amp_sp), float(m.amp_cp)

    def tab(kind, rows):
        cm, pr = np.zeros((rows, k), np.int32), np.zeros((rows, k), np.float32)
        rc = L.t2b200_pilot_tables(C.byref(m), kind, cm.ctypes.data, pr.ctypes.data)
        if rc != OK:
            raise T2Error('t2b200_pilot_tables rc=%d' % rc)
        return cm, pr
    t['data_map'], t['data_ref'] = tab(1, nd)
    cm, pr = tab(0, 1)
    t['p2_map'], t['p2_ref'] = cm[0], pr[0]
    if m.l_fc:
        cm, pr = tab(2, 1)
        t['fc_map'], t['fc_ref'] = cm[0], pr[0]
    for name, n in (('p2', m.c_p2), ('data', m.c_data), ('fc', m.n_fc)):
        if n:
            t['h_even_' + name], t['h_odd_' + name] = freq_deinterleaver_table(m.fft_size, n)
    return t


COMM_ID_BYTES = 128


def comm_unique_id():
    """the NCCL rendezvous id (rank 0 calls this and distributes the bytes)"""
    buf = (C.c_char * COMM_ID_BYTES)()
    rc = lib().t2b200_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != OK:
        raise T2Error('t2b200_comm_unique_id rc=%d (NCCL not loadable?)' % rc)
    return bytes(buf)


def cell_permutation(n_fec_blocks, cells_per_fec):
    out = np.zeros(n_fec_blocks * cells_per_fec, np.int32)
    rc = lib().t2b200_cell_permutation(n_fec_blocks, cells_per_fec, out.ctypes.data)
    if rc != OK:
        raise T2Error('t2b200_cell_permutation rc=%d' % rc)
    return out


def demap_address_table(fec_type, mod, code_rate):
    out = np.zeros(64800 if fec_type else 16200, np.int32)
    rc = lib().t2b200_demap_address_table(fec_type, mod, code_rate, out.ctypes.data)
    if rc != OK:
        raise T2Error('t2b200_demap_address_table rc=%d' % rc)
    return out


def freq_deinterleaver_table(fft_size, n_cells):
    """(h_even, h_odd) int32[n_cells] of address_freq_deinterleaver for a symbol kind with n_cells cells"""
    e, o = np.zeros(n_cells, np.int32), np.zeros(n_cells, np.int32)
    rc = lib().t2b200_freq_deinterleaver_table(fft_size, n_cells, e.ctypes.data, o.ctypes.data)
    if rc != OK:
        raise T2Error('t2b200_freq_deinterleaver_table rc=%d' % rc)
    return e, o
