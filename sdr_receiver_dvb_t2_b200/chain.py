"""Replay-mode receive chain over the C-ABI: FFT -> equalise / frequency de-interleave -> time / cell
de-interleave -> soft demap -> LDPC (+ BCH strip, BB descramble), whole T2 frames per call, everything
resident on the device between stages.  This is the throughput ("teacher-forced") mode of SURVEY 7.3-7:
the caller supplies per-symbol FFT windows (dvbt2_demodulator.cpp:332) of already synchronised frames.

Only orchestration lives here: every stage is a t2b200_* call, i.e. CUDA kernels of libt2b200.so.
"""
import numpy as np

from . import engine as E


class FrameChain:
    """One PLP (type 1, P_I = 1) in one transmission mode.

    tables: dict with the init-time tables the reference's pilot_generator / address_freq_deinterleaver hold
            (keys as oracle RefRx.tables() / tools.make_golden_tables.load()) and the mode parameters under 'p'.
    """

    def __init__(self, eng, tables, mod, cod, fec_type, n_blocks, ti_len, rotation=1, l1_post_size=360, plp=0, cell_offset=0):
        import torch
        self.torch = torch
        self.eng, self.t, self.p = eng, tables, tables['p']
        # the glue between the stages is torch copies: the engine must run on the stream they run on (calls are
        # asynchronous).  Call the chain under the torch stream that was current here.
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        p = self.p
        self.mod, self.cod, self.fec_type, self.rot, self.plp = mod, cod, fec_type, rotation, plp
        self.nbits = 64800 if fec_type else 16200
        self.cpf = self.nbits // (2 * (mod + 1))
        base = n_blocks // ti_len
        self.blocks = [base + (1 if j >= ti_len - n_blocks % ti_len else 0) for j in range(ti_len)]   # time_deinterleaver.cpp:275-282
        self.n_blocks = n_blocks
        # first PLP cell of the frame: behind the L1 cells (time_deinterleaver.cpp:44) and the cells of the PLPs in front
        # of this one (type 1, contiguous: l1_postsignalling_dynamic.start)
        self.p2_start = 1840 + l1_post_size + cell_offset
        self.n_data_sym = p['len_frame'] - p['n_p2'] - p['l_fc']
        self.code = eng.ldpc_code_id(fec_type, cod)
        t = tables
        eng.eq_configure(E_SYM_P2, 0, p['fft_size'], p['k_total'], p['l_nulls'], p['c_p2'], t['p2_map'][None], t['p2_ref'][None],
                         t['h_even_p2'], t['h_odd_p2'], t['amp_p2'])
        eng.eq_configure(E_SYM_DATA, p['n_p2'], p['fft_size'], p['k_total'], p['l_nulls'], p['c_data'], t['data_map'], t['data_ref'],
                         t['h_even_data'], t['h_odd_data'], t['amp_sp'], t['amp_cp'])
        if p['l_fc']:
            eng.eq_configure(E_SYM_FC, p['len_frame'] - 1, p['fft_size'], p['k_total'], p['l_nulls'], p['n_fc'], t['fc_map'][None],
                             t['fc_ref'][None], t['h_even_fc'], t['h_odd_fc'], t['amp_sp'])
        eng.ti_configure(plp, fec_type, mod, max(self.blocks))
        self.need = n_blocks * self.cpf
        # the same chain as ONE C call (t2b200_frames_decode): no glue copies, nothing of this class on the data path
        eng.frames_configure(fft_size=p['fft_size'], len_frame=p['len_frame'], n_p2=p['n_p2'], l_fc=p['l_fc'], c_p2=p['c_p2'],
                             c_data=p['c_data'], n_fc=p['n_fc'] if p['l_fc'] else 0, first_cell=self.p2_start, plp=plp, mod=mod,
                             rotation=rotation, fec_type=fec_type, code_rate=cod, n_blocks=n_blocks, ti_len=ti_len)
        self._bufs = {}
        self.events = None          # set to {} to collect per-stage CUDA events: name -> [(start, stop), ...]

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            b = self.torch.empty(shape, dtype=dtype, device='cuda:%d' % self.eng.device)
            self._bufs[key] = b
        return b

    def _timed(self, name):
        chain = self

        class _T:
            def __enter__(self_inner):
                if chain.events is not None:
                    st = chain.torch.cuda.current_stream()
                    self_inner.a = chain.torch.cuda.Event(enable_timing=True)
                    self_inner.a.record(st)

            def __exit__(self_inner, *exc):
                if chain.events is not None:
                    b = chain.torch.cuda.Event(enable_timing=True)
                    b.record(chain.torch.cuda.current_stream())
                    chain.events.setdefault(name, []).append((self_inner.a, b))
        return _T()

    def stage_ms(self):
        """median device milliseconds per stage from the collected events (a stage that had to wait for an allocation once
        does not move it)"""
        out = {}
        for k, v in (self.events or {}).items():
            ms = sorted(a.elapsed_time(b) for a, b in v)
            out[k] = ms[len(ms) // 2]
        return out

    def demodulate(self, time):
        """time: torch complex64 cuda [F][len_frame][fft_size] -> PLP cell stream [F][n_blocks*cpf] (arrival order),
        sro / phase feedback [F][len_frame]"""
        torch, p, eng = self.torch, self.p, self.eng
        F = time.shape[0]
        L, N = p['len_frame'], p['fft_size']
        freq = self._buf('freq', (F * L, N), torch.complex64)
        with self._timed('fft'):
            eng.fft(time.reshape(F * L, N), out=freq)
        freq3 = freq.reshape(F, L, N)
        # frame cell stream: P2 cells | data symbols | FC
        per_frame = p['c_p2'] + self.n_data_sym * p['c_data'] + (p['n_fc'] if p['l_fc'] else 0)
        cells = self._buf('cells', (F, per_frame), torch.complex64)
        sro = torch.zeros((F, L), dtype=torch.float32, device=cells.device)
        ph = torch.zeros((F, L), dtype=torch.float32, device=cells.device)
        # P2 symbols of all frames in one launch, data symbols of all frames in one launch (strided views are
        # materialised once: the equaliser wants [n][fft_size] / writes [n][n_out] contiguous)
        p2f = freq3[:, 0, :].contiguous()
        with self._timed('equalize_p2'):
            c, s, h = eng.equalize(E_SYM_P2, np.zeros(F, np.int32), p2f)
        cells[:, :p['c_p2']] = c
        sro[:, 0], ph[:, 0] = s, h
        nd = self.n_data_sym
        dfreq = freq3[:, p['n_p2']:p['n_p2'] + nd, :].reshape(F * nd, N)
        idx = np.tile(np.arange(p['n_p2'], p['n_p2'] + nd, dtype=np.int32), F)
        dfreq = dfreq.contiguous() if not dfreq.is_contiguous() else dfreq
        with self._timed('equalize_data'):
            c, s, h = eng.equalize(E_SYM_DATA, idx, dfreq)
        cells[:, p['c_p2']:p['c_p2'] + nd * p['c_data']] = c.reshape(F, nd * p['c_data'])
        sro[:, p['n_p2']:p['n_p2'] + nd], ph[:, p['n_p2']:p['n_p2'] + nd] = s.reshape(F, nd), h.reshape(F, nd)
        if p['l_fc']:
            fcf = freq3[:, L - 1, :].contiguous()
            c, s, h = eng.equalize(E_SYM_FC, np.full(F, L - 1, np.int32), fcf)
            cells[:, p['c_p2'] + nd * p['c_data']:] = c
            sro[:, L - 1], ph[:, L - 1] = s, h
        self.last_cells = cells
        return self.plp_cells(cells), sro, ph

    def plp_cells(self, cells):
        """this PLP's cells out of the frame cell streams [F][c_p2 + n_data*c_data (+ n_fc)] (another chain of the same
        frame geometry may have produced them: `other.last_cells`)"""
        return cells[:, self.p2_start:self.p2_start + self.need].contiguous()

    def fec(self, stream, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE, precision_in=None, want_llr=False, max_trials=25,
            want_status=True):
        """stream [F][n_blocks*cpf] complex64 cuda -> dict(bits [F*n_blocks][K_bch|K], trials_left, iterations, snr, llr?)"""
        F = stream.shape[0]
        blocks = self.blocks * F
        with self._timed('ti_deinterleave'):
            ti = self.eng.ti_deinterleave(self.plp, stream.reshape(-1), blocks)
        with self._timed('demap'):
            d = self.eng.demap(ti, blocks, self.mod, self.rot, self.fec_type, self.cod, precision_in=precision_in)
        with self._timed('ldpc_bch'):
            r = self.eng.ldpc_decode(self.code, d['llr'], flags=flags, max_trials=max_trials, want_status=want_status)
        r['snr'], r['precision'] = d['snr'], d['precision']
        if want_llr:
            r['llr'], r['ti'] = d['llr'], ti
        return r

    def decode_frames_fused(self, time, flags=E.LDPC_GROUP32 | E.LDPC_BCH_DESCRAMBLE, max_trials=25, want_status=True, out=None,
                            scale=None):
        """the whole chain through t2b200_frames_decode (the engine keeps ONE frame configuration: the last chain built);
        scale given: `time` holds int16 (I, Q) pairs [F][len_frame][fft_size][2] (t2b200_frames_decode_i16)"""
        return self.eng.frames_decode(time, flags=flags, max_trials=max_trials, want_status=want_status, out=out, scale=scale)

    def decode_frames(self, time, host_feedback=True, **kw):
        """host_feedback=False leaves sro / phase / snr / precision on the device: nothing in the call waits for the GPU"""
        stream, sro, ph = self.demodulate(time)
        r = self.fec(stream, **kw)
        r['sro'], r['phase'] = sro, ph
        if host_feedback:
            self.eng.sync()                                   # the engine's stream need not be torch's current stream
            for k in ('sro', 'phase', 'snr', 'precision'):
                r[k] = r[k].cpu().numpy()
        return r


E_SYM_P2, E_SYM_DATA, E_SYM_FC = 0, 1, 2
