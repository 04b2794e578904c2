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


_bch_gen = {}


def bch_generator(short_frame, t):
    """generator polynomial of the DVB-T2 outer code as a Python int (bit i = coefficient of x^i): product of the minimal
    polynomials of alpha, alpha^3, ..., alpha^(2t-1), alpha a root of the primitive polynomial of EN 302 755 table 6a / 6b"""
    key = (bool(short_frame), t)
    if key not in _bch_gen:
        m = 14 if short_frame else 16
        prim = (1 << 14 | 1 << 5 | 1 << 3 | 1 << 1 | 1) if short_frame else (1 << 16 | 1 << 5 | 1 << 3 | 1 << 2 | 1)
        n = (1 << m) - 1
        ex, lg, x = [0] * (2 * n), [0] * (n + 1), 1
        for i in range(n):
            ex[i], lg[x] = x, i
            x <<= 1
            if x >> m:
                x ^= prim
        for i in range(n, 2 * n):
            ex[i] = ex[i - n]
        gen = 1
        for i in range(1, t + 1):
            e = (2 * i - 1) % n
            conj, c = [], e
            while c not in conj:
                conj.append(c)
                c = (c * 2) % n
            poly = [1]
            for c in conj:
                a, new = ex[c], [0] * (len(poly) + 1)
                for k, pk in enumerate(poly):
                    new[k + 1] ^= pk
                    if pk:
                        new[k] ^= ex[lg[pk] + lg[a]]
                poly = new
            mp = sum((pk & 1) << k for k, pk in enumerate(poly))
            prod, a, sh = 0, gen, 0                          # carry-less multiply gen * mp
            while mp >> sh:
                if (mp >> sh) & 1:
                    prod ^= a << sh
                sh += 1
            gen = prod
        _bch_gen[key] = (gen, m * t)
    return _bch_gen[key]


def bch_parity(bits, short_frame, t):
    """bits uint8[k] (first bit = highest power) -> parity uint8[m*t]: remainder of x^(m t) * msg(x) modulo g(x), MSB first"""
    gen, deg = bch_generator(short_frame, t)
    mask = (1 << deg) - 1
    table = []
    for b in range(256):                                        # (b << deg) mod g, one message byte at a time
        r = b << deg
        for k in range(deg + 7, deg - 1, -1):
            if (r >> k) & 1:
                r ^= gen << (k - deg)
        table.append(r & mask)
    pad = (-len(bits)) % 8
    by = np.packbits(np.concatenate([np.zeros(pad, np.uint8), np.asarray(bits, np.uint8)]))
    reg = 0
    for b in by.tolist():
        reg = ((reg << 8) & mask) ^ table[(reg >> (deg - 8)) ^ b]
    return np.array([(reg >> (deg - 1 - k)) & 1 for k in range(deg)], np.uint8)


def bch_t(code):
    return 10 if code in (2, 5) else 12


def make_bbframes(code, n, rng, bch=False):
    """n BBFRAMEs of K_bch bits, high-efficiency mode, carrying 187-byte packets of random payload.
    Returns (descrambled bits [n][K_bch] -- what the receiver must output --, scrambled+padded info [n][K_ldpc]).
    bch: fill the BCH parity bits [K_bch, K_ldpc) (EN 302 755 6.1.1) instead of leaving them zero (the reference never reads
    them, bch_decoder.cpp:136)."""
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
    info[:, :kb] = scr                                           # BCH parity bits [K_bch, K_ldpc) left zero unless asked for
    if bch:
        for i in range(n):
            info[i, kb:] = bch_parity(scr[i], code >= 6, bch_t(code))
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

    def __init__(self, tables, mod, cod, fec_normal, n_blocks, ti_len, rotation=True, l1_post_size=360, seed=1, bch=False):
        from sdr_receiver_dvb_t2_b200 import engine as E
        self.t, self.p = tables, tables['p']
        self.bch = bch
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
        bb, info = make_bbframes(self.code, self.nb, self.rng, bch=self.bch)
        cw = ldpc_encode(self.code, info)
        cells = self.fec_cells(cw)
        stream, off = [], 0
        for nf in self.blocks:
            stream.append(self.ti_stream(cells[off:off + nf]))
            off += nf
        return bb, cw, cells, np.concatenate(stream)

    # ---- one T2 frame ----
    def frame(self, noise_cn_db=None, scale=200.0, extra_streams=(), l1_cells=None):
        """-> dict(time complex64[len_frame][fft_size], bb bits [n_blocks][K_bch], cells ...).
        extra_streams: cell streams of further (type-1, contiguous) PLPs placed right behind this one."""
        p, t = self.p, self.t
        bb, cw, cells, stream = self.plp_stream()
        # frame cell stream: L1 cells (BPSK +-1, never parsed in replay mode) | PLP(s) | dummy cells
        cap = p['c_p2'] - self.p2_start + self.n_data_sym * p['c_data'] + (p['n_fc'] if p['l_fc'] else 0)
        plps = np.concatenate([stream] + list(extra_streams))
        assert len(plps) <= cap, 'PLPs do not fit the frame'
        dummy = qam_map(self.rng.integers(0, 2, (cap - len(plps), self.bpc), dtype=np.uint8), self.mod)
        l1 = (1.0 - 2.0 * self.rng.integers(0, 2, self.p2_start)).astype(np.complex128) if l1_cells is None else l1_cells
        assert len(l1) == self.p2_start
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


# =============================================================================================
# Whole-signal transmitter: P1 preamble + L1-pre / L1-post signalling + guard intervals -> the int16 I/Q stream a device
# front-end hands to dvbt2_demodulator::execute (dvbt2_demodulator.cpp:145).  Test-only, like everything in this file.
P1_ACTIVE = [
    44, 45, 47, 51, 54, 59, 62, 64, 65, 66, 70, 75, 78, 80, 81, 82, 84, 85, 87, 88, 89, 90, 94, 96, 97, 98, 102, 107, 110, 112,
    113, 114, 116, 117, 119, 120, 121, 122, 124, 125, 127, 131, 132, 133, 135, 136, 137, 138, 142, 144, 145, 146, 148, 149, 151,
    152, 153, 154, 158, 160, 161, 162, 166, 171, 172, 173, 175, 179, 182, 187, 190, 192, 193, 194, 198, 203, 206, 208, 209, 210,
    212, 213, 215, 216, 217, 218, 222, 224, 225, 226, 230, 235, 238, 240, 241, 242, 244, 245, 247, 248, 249, 250, 252, 253, 255,
    259, 260, 261, 263, 264, 265, 266, 270, 272, 273, 274, 276, 277, 279, 280, 281, 282, 286, 288, 289, 290, 294, 299, 300, 301,
    303, 307, 310, 315, 318, 320, 321, 322, 326, 331, 334, 336, 337, 338, 340, 341, 343, 344, 345, 346, 350, 352, 353, 354, 358,
    363, 364, 365, 367, 371, 374, 379, 382, 384, 385, 386, 390, 395, 396, 397, 399, 403, 406, 411, 412, 413, 415, 419, 420, 421,
    423, 424, 425, 426, 428, 429, 431, 435, 438, 443, 446, 448, 449, 450, 454, 459, 462, 464, 465, 466, 468, 469, 471, 472, 473,
    474, 478, 480, 481, 482, 486, 491, 494, 496, 497, 498, 500, 501, 503, 504, 505, 506, 508, 509, 511, 515, 516, 517, 519, 520,
    521, 522, 526, 528, 529, 530, 532, 533, 535, 536, 537, 538, 542, 544, 545, 546, 550, 555, 558, 560, 561, 562, 564, 565, 567,
    568, 569, 570, 572, 573, 575, 579, 580, 581, 583, 584, 585, 586, 588, 589, 591, 595, 598, 603, 604, 605, 607, 611, 612, 613,
    615, 616, 617, 618, 622, 624, 625, 626, 628, 629, 631, 632, 633, 634, 636, 637, 639, 643, 644, 645, 647, 648, 649, 650, 654,
    656, 657, 658, 660, 661, 663, 664, 665, 666, 670, 672, 673, 674, 678, 683, 684, 689, 692, 696, 698, 699, 701, 702, 703, 704,
    706, 707, 708, 712, 714, 715, 717, 718, 719, 720, 722, 723, 725, 726, 727, 729, 733, 734, 735, 736, 738, 739, 740, 744, 746,
    747, 748, 753, 756, 760, 762, 763, 765, 766, 767, 768, 770, 771, 772, 776, 778, 779, 780, 785, 788, 792, 794, 795, 796, 801,
    805, 806, 807, 809]                                  # EN 302 755 table 61 (p1_symbol.h:60-110)
# EN 302 755 table 60: the S1 (3 bit -> 64 chip) and S2 (4 bit -> 256 chip) modulation signalling sequences, hex
P1_S1 = ['124721741D482E7B', '47127421481D7B2E', '217412472E7B1D48', '742147127B2E481D',
         '1D482E7B12472174', '481D7B2E47127421', '2E7B1D4821741247', '7B2E481D74214712']
P1_S2 = ['121D4748212E747B1D1248472E217B7412E247B721D174841DED48B82EDE7B8B',
         '4748121D747B212E48471D127B742E2147B712E2748421D148B81DED7B8B2EDE',
         '212E747B121D47482E217B741D12484721D1748412E247B72EDE7B8B1DED48B8',
         '747B212E4748121D7B742E2148471D12748421D147B712E27B8B2EDE48B81DED',
         '1D1248472E217B74121D4748212E747B1DED48B82EDE7B8B12E247B721D17484',
         '48471D127B742E214748121D747B212E48B81DED7B8B2EDE47B712E2748421D1',
         '2E217B741D124847212E747B121D47482EDE7B8B1DED48B821D1748412E247B7',
         '7B742E2148471D12747B212E4748121D7B8B2EDE48B81DED748421D147B712E2',
         '12E247B721D174841DED48B82EDE7B8B121D4748212E747B1D1248472E217B74',
         '47B712E2748421D148B81DED7B8B2EDE4748121D747B212E48471D127B742E21',
         '21D1748412E247B72EDE7B8B1DED48B8212E747B121D47482E217B741D124847',
         '748421D147B712E27B8B2EDE48B81DED747B212E4748121D7B742E2148471D12',
         '1DED48B82EDE7B8B12E247B721D174841D1248472E217B74121D4748212E747B',
         '48B81DED7B8B2EDE47B712E2748421D148471D127B742E214748121D747B212E',
         '2EDE7B8B1DED48B821D1748412E247B72E217B741D124847212E747B121D4748',
         '7B8B2EDE48B81DED748421D147B712E27B742E2148471D12747B212E4748121D']


def _hex_bits(h):
    return np.array([(int(c, 16) >> (3 - k)) & 1 for c in h for k in range(4)], np.uint8)


def p1_symbol(s1, s2):
    """EN 302 755 9.8: 2048 samples C | A | B, unit mean power."""
    mss = np.concatenate([_hex_bits(P1_S1[s1]), _hex_bits(P1_S2[s2]), _hex_bits(P1_S1[s1])])
    dif = np.cumprod(1 - 2 * mss.astype(np.int64))                 # DBPSK, reference symbol +1
    sr, scr = 0x4e46, np.zeros(384, np.int64)                      # PRBS of 9.8.2.3 as p1_symbol.cpp:46-55 runs it
    for i in range(384):
        b = (sr ^ (sr >> 1)) & 1
        scr[i] = -1 if b else 1
        sr >>= 1
        if b:
            sr |= 0x4000
    spec = np.zeros(1024, np.complex128)
    spec[(np.array(P1_ACTIVE) - 426) % 1024] = dif * scr
    a = np.fft.ifft(spec) * 1024 / np.sqrt(384)
    n = np.arange(2048)
    sh = np.exp(2j * np.pi * n / 1024)
    return np.concatenate([a[:542] * sh[:542], a, a[542:] * sh[1566:]])


def crc32_bits(bits):
    """CRC-32 (0x04C11DB7, all-ones start, no final inversion) over a bit array, as p2_symbol.cpp:308-318 checks it"""
    crc = 0xffffffff
    for b in bits:
        fb = int(b) ^ ((crc >> 31) & 1)
        crc = (crc << 1) & 0xffffffff
        if fb:
            crc ^= 0x04C11DB7
    return crc


def _field(v, n):
    return [(int(v) >> (n - 1 - k)) & 1 for k in range(n)]


def l1_pre_bits(m, l1_post_mod, l1_post_size, l1_post_info_size, s2_field1):
    """EN 302 755 7.2.2: the 168 signalling bits + CRC-32 (all p2_symbol::l1_pre_info reads, p2_symbol.cpp:301-500)"""
    f = []
    for v, n in ((0, 8), (m.carrier_mode, 1), (0, 3), (s2_field1, 3), (0, 1), (0, 1), (m.guard_interval_mode, 3), (m.papr_mode, 4),
                 (l1_post_mod, 4), (0, 2), (0, 2), (l1_post_size, 18), (l1_post_info_size, 18), (m.pilot_pattern, 4), (0, 8),
                 (0, 16), (0x3085, 16), (0x8001, 16), (2, 8), (m.n_data, 12), (0, 3), (0, 1), (1, 3), (0, 3), (0, 4), (0, 1),
                 (0, 1), (0, 4)):
        f += _field(v, n)
    assert len(f) == 168
    return np.array(f + _field(crc32_bits(f), 32), np.uint8)


def l1_post_bits(plps, frame_idx):
    """EN 302 755 7.2.3: configurable + dynamic L1-post for NUM_RF = 1, no FEF, no auxiliary streams, + CRC-32, with the
    field offsets p2_symbol::l1_post_info and its parsers use (p2_symbol.cpp:671-1005).
    plps: dicts(id, cod, mod, rot, fec, blocks_max, ti_len, ti_type, start, num_blocks)"""
    f = _field(0, 15) + _field(len(plps), 8) + _field(0, 4) + _field(0, 8)                 # SUB_SLICES, NUM_PLP, NUM_AUX, AUX_RFU
    f += _field(0, 3) + _field(666000000, 32)                                              # RF_IDX, FREQUENCY
    for p in plps:
        for v, n in ((p['id'], 8), (1, 3), (3, 5), (0, 1), (0, 3), (0, 8), (p['id'], 8), (p['cod'], 3), (p['mod'], 3),
                     (p['rot'], 1), (p['fec'], 2), (p['blocks_max'], 10), (1, 8), (p['ti_len'], 8), (p['ti_type'], 1),
                     (0, 1), (0, 1), (0, 11), (1, 2), (0, 1), (0, 1)):
            f += _field(v, n)
    f += _field(0, 32)                                                                     # FEF_LENGTH_MSB, RESERVED_2
    assert len(f) == 70 + 89 * len(plps) + 32
    f += _field(frame_idx, 8) + _field(0, 22) + _field(0, 22) + _field(0, 8) + _field(0, 3) + _field(0, 8)
    for p in plps:
        f += _field(p['id'], 8) + _field(p['start'], 22) + _field(p['num_blocks'], 10) + _field(0, 8)
    f += _field(0, 8)
    return np.array(f + _field(crc32_bits(f), 32), np.uint8)


class Transmitter:
    """T2 frames of one PLP as the int16 I/Q sample stream of an 8 MHz channel sampled at 64/7 MHz: P1, P2 carrying real
    L1-pre / L1-post signalling (QPSK L1-post; only the systematic bits + CRC are meaningful, which is all the reference
    reads -- it never runs the L1 LDPC / BCH), data symbols (+ frame closing), guard intervals.
    m: engine.Mode; the carrier / pilot tables are the natively built ones (engine.mode_tables)."""

    def __init__(self, m, mod, cod, fec_normal, n_blocks, ti_len, rotation=True, seed=1, rms=3500.0):
        from sdr_receiver_dvb_t2_b200 import engine as E
        self.m = m
        self.tables = E.mode_tables(m)
        self.l1_post_mod = 1                                         # QPSK (BPSK would overflow the reference's bit buffers, p2_symbol.cpp:383-386)
        self.plp = dict(id=0, cod=cod, mod=mod, rot=int(rotation), fec=int(fec_normal), blocks_max=n_blocks, ti_len=ti_len,
                        ti_type=0, start=0, num_blocks=n_blocks)
        self.info_size = len(l1_post_bits([self.plp], 0)) - 32
        self.l1_post_size = 750                                      # cells: EN 302 755 7.3.1.2 for K_sig = 350, QPSK, one P2 symbol
        assert self.info_size + 32 <= 2 * self.l1_post_size
        self.mod = Modulator(self.tables, mod, cod, fec_normal, n_blocks, ti_len, rotation=rotation,
                             l1_post_size=self.l1_post_size, seed=seed)
        self.rng = np.random.default_rng(seed + 7777)
        self.rms = rms
        self.s2 = {16384: 4, 32768: 5}[m.fft_size] << 1              # S2 field 1 (FFT size), field 2 = 0 (not mixed)
        self.p1 = p1_symbol(0, self.s2)                              # S1 = 000: T2 SISO
        self.frame_idx = 0
        self._scale = None

    def l1_cells(self):
        pre = l1_pre_bits(self.m, self.l1_post_mod, self.l1_post_size, self.info_size, self.s2 >> 1)
        pre = np.concatenate([pre, self.rng.integers(0, 2, 1840 - len(pre), dtype=np.uint8)])      # BCH / LDPC parity: never read
        post = l1_post_bits([self.plp], self.frame_idx & 0xff)
        post = np.concatenate([post, self.rng.integers(0, 2, 2 * self.l1_post_size - len(post), dtype=np.uint8)])
        qp = ((1.0 - 2.0 * post[0::2]) + 1j * (1.0 - 2.0 * post[1::2])) / np.sqrt(2.0)
        return np.concatenate([(1.0 - 2.0 * pre).astype(np.complex128), qp])

    def frame(self):
        """-> (samples complex128 [2048 + len_frame * (N + GI)], unit-ish power; modulator frame dict)"""
        f = self.mod.frame(noise_cn_db=None, scale=1.0, l1_cells=self.l1_cells())
        self.frame_idx += 1
        N, gi = self.m.fft_size, self.m.guard_interval_size
        t = f['time'].astype(np.complex128) * N / np.sqrt(self.m.k_total)          # ~unit power per sample
        sym = np.concatenate([t[:, N - gi:], t], axis=1).reshape(-1)
        return np.concatenate([self.p1, sym]), f

    def stream(self, n_frames, cn_db=None, lead=4096):
        """-> (I int16, Q int16, list of modulator frame dicts).  `lead` noise-only samples in front."""
        parts, frames = [], []
        for _ in range(n_frames):
            s, f = self.frame()
            parts.append(s)
            frames.append(f)
        x = np.concatenate([np.zeros(lead, np.complex128)] + parts + [np.zeros(lead, np.complex128)])
        p_sig = np.mean(np.abs(np.concatenate(parts)) ** 2)
        if cn_db is not None:
            sig = np.sqrt(p_sig * 10 ** (-cn_db / 10) / 2)
            x = x + sig * (self.rng.standard_normal(len(x)) + 1j * self.rng.standard_normal(len(x)))
        x *= self.rms * np.sqrt(2.0) / np.sqrt(p_sig)
        i16 = np.clip(np.rint(x.real), -32768, 32767).astype(np.int16)
        q16 = np.clip(np.rint(x.imag), -32768, 32767).astype(np.int16)
        return i16, q16, frames
