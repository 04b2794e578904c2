"""Test-only DVB-T2 modulator (the reference has no transmitter).

Builds T2 frames the receive chain of this repo -- and the reference's -- can decode: random BBFRAME
payload -> BB scrambling -> (BCH parity left zero: the reference discards it, bch_decoder.cpp:136) ->
LDPC encoding -> bit interleaving + demux -> rotated-QAM mapping with cyclic Q delay -> cell and time
interleaving -> P2 / data-symbol cell mapping through the frequency interleaver -> pilots -> IFFT.
Every permutation is taken from the RECEIVER's tables and applied backwards, so the modulator is by
construction the inverse of the receive path under test (conventions: SURVEY appendix A).

Output is per-OFDM-symbol time-domain buffers (the `in_fft` of dvbt2_demodulator.cpp:332, i.e. after
synchronisation / guard-interval removal): the replay ("teacher-forced") mode of SURVEY 7.3-7.
Pure numpy/scipy; the only thing taken from the product library are its host-side table builders.
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NORM = [0.707106781, 0.316227766, 0.15430335, 0.076696499]
ROT = [0.506145483, 0.293215314, 0.150098316, 0.062418810]
K_BCH = {0: 32208, 1: 38688, 2: 43040, 3: 48408, 4: 51648, 5: 53840, 6: 7032, 7: 9552, 8: 10632, 9: 11712, 10: 12432, 11: 13152}

_codes = None


def ldpc_tables():
    """parse the product's generated table file -> {name: (N, K, rows)}"""
    global _codes
    if _codes is None:
        txt = open(os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'csrc', 'ldpc_tables_data.inc')).read()
        deg = {m.group(1): [int(x) for x in m.group(2).split(',')] for m in
               re.finditer(r'kRowDeg_(\w+)\[\d+\] = \{([^}]*)\}', txt)}
        addr = {m.group(1): [int(x) for x in m.group(2).replace('\n', ' ').split(',') if x.strip()] for m in
                re.finditer(r'kAddr_(\w+)\[\d+\] = \{([^}]*)\}', txt)}
        order = re.findall(r'\{"(\w+)", (\d+), (\d+), (\d+),', txt)
        _codes = []
        for name, n, k, nrows in order:
            rows, p = [], 0
            for d in deg[name]:
                rows.append(addr[name][p:p + d])
                p += d
            _codes.append((int(n), int(k), rows))
    return _codes


_enc = {}


def ldpc_encode(code, info):
    """info uint8[n][K] -> codewords uint8[n][N] in the order the receiver consumes them (parity interleaved)"""
    import scipy.sparse as sp
    N, K, rows = ldpc_tables()[code]
    R = N - K
    q = R // 360
    if code not in _enc:
        ri, ci = [], []
        for g, row in enumerate(rows):
            m = np.arange(360)
            for x in row:
                ri.append((x + q * m) % R)
                ci.append(360 * g + m)
        ri, ci = np.concatenate(ri), np.concatenate(ci)
        _enc[code] = sp.csr_matrix((np.ones(len(ri), np.int32), (ri, ci)), shape=(R, K))
    A = _enc[code]
    info = np.ascontiguousarray(info, np.uint8).reshape(-1, K)
    acc = (A @ info.T.astype(np.int32)) & 1                    # [R][n]
    p = np.bitwise_xor.accumulate(acc.astype(np.uint8), axis=0)  # p[i] ^= p[i-1]
    par = p.reshape(360, q, -1).transpose(1, 0, 2).reshape(R, -1)   # u[K + 360 t + s] = p[q s + t]
    return np.concatenate([info, par.T.astype(np.uint8)], axis=1)


def bb_prbs(n):
    out = np.zeros(n, np.uint8)
    sr = 0x4A80
    for i in range(n):
        b = (sr ^ (sr >> 1)) & 1
        out[i] = b
        sr >>= 1
        if b:
            sr |= 0x4000
    return out


def crc8_bits(bits):
    """CRC-8 (poly 0xD5) over a bit array, MSB first, as EN 302 755 annex F"""
    crc = 0
    for b in bits:
        fb = ((crc >> 7) & 1) ^ int(b)
        crc = (crc << 1) & 0xff
        if fb:
            crc ^= 0xD5
    return crc


def make_bbframes(code, n, rng):
    """n BBFRAMEs of K_bch bits, high-efficiency mode, carrying 187-byte packets of random payload.
    Returns (descrambled bits [n][K_bch] -- what the receiver must output --, scrambled+padded info [n][K_ldpc])."""
    N, K, _ = ldpc_tables()[code]
    kb = K_BCH[code]
    dfl = ((kb - 80) // 8) * 8
    frames = np.zeros((n, kb), np.uint8)
    syncd = 0
    for i in range(n):
        hdr = np.zeros(80, np.uint8)
        # MATYPE-1: TS (11), SIS (1), CCM (1), ISSYI 0, NPD 0, EXT 00 ; MATYPE-2 0
        hdr[0:8] = [1, 1, 1, 1, 0, 0, 0, 0]
        hdr[16:32] = 0                                           # ISSY / UPL field unused in HEM
        hdr[32:48] = [(dfl >> (15 - b)) & 1 for b in range(16)]
        hdr[48:56] = 0
        hdr[56:72] = [(syncd >> (15 - b)) & 1 for b in range(16)]
        c = crc8_bits(hdr[:72]) ^ 1                              # HEM: CRC-8 XOR MODE (1)
        hdr[72:80] = [(c >> (7 - b)) & 1 for b in range(8)]
        frames[i, :80] = hdr
        frames[i, 80:80 + dfl] = rng.integers(0, 2, dfl, dtype=np.uint8)
        syncd = (syncd - dfl) % (187 * 8)
    scr = frames ^ bb_prbs(kb)[None, :]
    info = np.zeros((n, K), np.uint8)
    info[:, :kb] = scr                                           # BCH parity bits [K_bch, K_ldpc) left zero
    return frames, info


def qam_map(bits, mod):
    """bits uint8[n_cells][2*(mod+1)] in the demapper's production order (L0(I),L0(Q),L1(I),L1(Q),...) -> complex cells"""
    a = NORM[mod]
    nl = mod + 1

    def axis(b):                                                # b[:, l], l = 0..mod ; LLR > 0 <=> bit 0
        mag = np.full(len(b), 1.0)
        for l in range(nl - 1, 0, -1):                          # innermost level first
            mag = (1 << (nl - l)) + np.where(b[:, l] == 0, 1.0, -1.0) * mag
        return np.where(b[:, 0] == 0, 1.0, -1.0) * mag * a
    return axis(bits[:, 0::2]) + 1j * axis(bits[:, 1::2])


class Modulator:
    """One PLP, type-1, in the geometry of a table fixture (tests/golden/tables_*.npz)."""

    def __init__(self, tables, mod, cod, fec_normal, n_blocks, ti_len, rotation=True, l1_post_size=360, seed=1):
        from sdr_receiver_dvb_t2_b200 import engine as E
        self.t, self.p = tables, tables['p']
        self.mod, self.cod, self.fec, self.nb, self.ti_len, self.rot = mod, cod, int(fec_normal), n_blocks, ti_len, rotation
        self.code = (0 if fec_normal else 6) + cod
        self.N = 64800 if fec_normal else 16200
        self.bpc = 2 * (mod + 1)
        self.cpf = self.N // self.bpc
        base = n_blocks // ti_len
        self.blocks = [base + (1 if j >= ti_len - n_blocks % ti_len else 0) for j in range(ti_len)]
        self.perm = E.cell_permutation(max(self.blocks), self.cpf)
        self.addr = E.demap_address_table(self.fec, mod, cod)
        self.p2_start = 1840 + l1_post_size
        self.rng = np.random.default_rng(seed)
        p = self.p
        self.n_data_sym = p['len_frame'] - p['n_p2'] - p['l_fc']
        cap = p['c_p2'] - self.p2_start + self.n_data_sym * p['c_data'] + (p['n_fc'] if p['l_fc'] else 0)
        assert n_blocks * self.cpf <= cap, 'PLP does not fit the frame'

    # ---- FEC blocks -> cells in arrival (time-interleaved) order ----
    def fec_cells(self, cw):
        """cw uint8[n][N] -> cells complex[n][cpf] as the receiver's deinterleaved TI block must hold them"""
        n = cw.shape[0]
        bits = cw[:, self.addr].reshape(n * self.cpf, self.bpc)
        c = qam_map(bits, self.mod).reshape(n, self.cpf)
        if self.rot:
            c = c * np.exp(1j * ROT[self.mod])
        return c

    def ti_stream(self, cells):
        """deinterleaved cells [n_fec][cpf] of ONE TI block -> arrival-order stream (cell + time interleaver, Q delay)"""
        n = cells.shape[0]
        flat = cells.reshape(-1)
        a = np.arange(n * self.cpf)
        qa = np.where(a % self.cpf == 0, a + self.cpf - 1, a - 1)
        tx = flat.real + 1j * flat.imag[qa]                      # address a carries (I_a, Q_{a-1 cyclic})
        rows, cols = self.cpf // 5, 5 * n
        k = np.arange(n * self.cpf)
        d = (k % cols) * rows + k // cols
        return tx[self.perm[:n * self.cpf][d]]

    # ---- one interleaving frame of this PLP: BBFRAMEs -> cells in arrival order ----
    def plp_stream(self):
        bb, info = make_bbframes(self.code, self.nb, self.rng)
        cw = ldpc_encode(self.code, info)
        cells = self.fec_cells(cw)
        stream, off = [], 0
        for nf in self.blocks:
            stream.append(self.ti_stream(cells[off:off + nf]))
            off += nf
        return bb, cw, cells, np.concatenate(stream)

    # ---- one T2 frame ----
    def frame(self, noise_cn_db=None, scale=200.0, extra_streams=()):
        """-> dict(time complex64[len_frame][fft_size], bb bits [n_blocks][K_bch], cells ...).
        extra_streams: cell streams of further (type-1, contiguous) PLPs placed right behind this one."""
        p, t = self.p, self.t
        bb, cw, cells, stream = self.plp_stream()
        # frame cell stream: L1 cells (BPSK +-1, never parsed in replay mode) | PLP(s) | dummy cells
        cap = p['c_p2'] - self.p2_start + self.n_data_sym * p['c_data'] + (p['n_fc'] if p['l_fc'] else 0)
        plps = np.concatenate([stream] + list(extra_streams))
        assert len(plps) <= cap, 'PLPs do not fit the frame'
        dummy = qam_map(self.rng.integers(0, 2, (cap - len(plps), self.bpc), dtype=np.uint8), self.mod)
        l1 = (1.0 - 2.0 * self.rng.integers(0, 2, self.p2_start)).astype(np.complex128)
        allc = np.concatenate([l1, plps, dummy])
        syms = []
        # P2 (idx_symbol 0 -> h_odd), data symbols (parity of idx), frame closing
        pos = 0
        layout = [(0, 0, t['p2_map'], t['p2_ref'], p['c_p2'])]
        for s in range(self.n_data_sym):
            layout.append((1, p['n_p2'] + s, t['data_map'][s], t['data_ref'][s], p['c_data']))
        if p['l_fc']:
            layout.append((2, p['len_frame'] - 1, t['fc_map'], t['fc_ref'], p['n_fc']))
        names = {0: 'p2', 1: 'data', 2: 'fc'}
        freq = np.zeros((len(layout), p['fft_size']), np.complex128)
        for i, (kind, idx, cmap, ref, nc) in enumerate(layout):
            h = t['h_odd_' + names[kind]] if idx % 2 == 0 else t['h_even_' + names[kind]]
            sc = allc[pos:pos + nc]
            pos += nc
            x = ref.astype(np.complex128)
            dmask = cmap == 1
            if kind == 0:
                dmask = dmask.copy()
                dmask[p['k_total'] // 2] = False                 # the reference never reads the P2 centre carrier
            nd = int(dmask.sum())
            x[dmask] = sc[h[:nd]]                                # dd-th data carrier carries cell h[dd]
            freq[i, p['l_nulls']:p['l_nulls'] + p['k_total']] = x * scale
        time = np.fft.ifft(np.fft.ifftshift(freq, axes=1), axis=1)
        if noise_cn_db is not None:
            # C/N over the active carriers: per-sample noise variance after the receiver's unnormalised FFT
            sig = scale * 10 ** (-noise_cn_db / 20) / np.sqrt(2) / np.sqrt(p['fft_size'])
            time = time + sig * (self.rng.standard_normal(time.shape) + 1j * self.rng.standard_normal(time.shape))
        return {'time': time.astype(np.complex64), 'bb': bb, 'cw': cw, 'cells': cells, 'stream': stream,
                'blocks': list(self.blocks)}
