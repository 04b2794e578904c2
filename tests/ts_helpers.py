"""BBFRAME streams for the TS re-packetiser tests: a continuous TS cut into BBFRAMEs (HEM or NM) plus fault cases."""
import numpy as np

from tools.modulator import crc8_bits


def _crc8_bytes(data):
    c = 0
    for b in data:
        for j in range(7, -1, -1):
            fb = ((c >> 7) & 1) ^ ((int(b) >> j) & 1)
            c = (c << 1) & 0xff
            if fb:
                c ^= 0xD5
    return c


def header(dfl, syncd, hem, upl=188 * 8, sync=0x47):
    hdr = np.zeros(80, np.uint8)
    hdr[0:8] = [1, 1, 1, 1, 0, 0, 0, 0]                       # TS, SIS, CCM
    hdr[16:32] = [(upl >> (15 - b)) & 1 for b in range(16)] if not hem else 0
    hdr[32:48] = [(dfl >> (15 - b)) & 1 for b in range(16)]
    hdr[48:56] = [(sync >> (7 - b)) & 1 for b in range(8)] if not hem else 0
    hdr[56:72] = [(syncd >> (15 - b)) & 1 for b in range(16)]
    c = crc8_bits(hdr[:72]) ^ (1 if hem else 0)
    hdr[72:80] = [(c >> (7 - b)) & 1 for b in range(8)]
    return hdr


def ts_stream(n_packets, rng):
    """188-byte packets: 0x47, PID / counter-like bytes, random payload"""
    p = rng.integers(0, 256, (n_packets, 188), dtype=np.uint8)
    p[:, 0] = 0x47
    p[:, 1] &= 0x1f                                            # transport_error_indicator clear
    return p


def bbframes(k_bch, dfl_bytes, n_frames, hem, rng, faults=()):
    """Cut a TS into n_frames BBFRAMEs of k_bch bits (one byte per bit).  HEM: sync bytes removed (187-byte UPs);
    NM: sync byte of each packet replaced by the CRC-8 of the previous packet.  dfl_bytes: int or list (per frame).
    faults: set of (frame, kind) with kind in {'crc', 'syncd65535', 'syncd_plus', 'syncd_minus'}.
    Returns (frames uint8[n][k_bch], ts packets uint8[m][188])."""
    dfls = [dfl_bytes] * n_frames if np.isscalar(dfl_bytes) else list(dfl_bytes)
    up = 187 if hem else 188
    total = sum(dfls)
    packets = ts_stream(total // up + 2, rng)
    if hem:
        air = packets[:, 1:].reshape(-1)
    else:
        air = packets.copy()
        for i in range(len(packets)):
            air[i, 0] = _crc8_bytes(packets[i - 1, 1:]) if i else 0
        air = air.reshape(-1)
    frames = np.zeros((n_frames, k_bch), np.uint8)
    pos = 0
    faults = dict(((f, k), True) for f, k in faults)
    for i, dfl_b in enumerate(dfls):
        syncd = ((-pos) % up) * 8                               # bits to the first UP that STARTS in this data field
        if (i, 'syncd65535') in faults:
            syncd = 65535
        if (i, 'syncd_plus') in faults:
            syncd += 16
        if (i, 'syncd_minus') in faults:
            syncd = max(0, syncd - 16)
        hdr = header(dfl_b * 8, syncd, hem)
        if (i, 'crc') in faults:
            hdr[79] ^= 1
        frames[i, :80] = hdr
        frames[i, 80:80 + 8 * dfl_b] = np.unpackbits(air[pos:pos + dfl_b])
        pos += dfl_b
    return frames, packets
