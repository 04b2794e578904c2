"""N1 on the GPU: t2b200_ts_packetize against the CPU oracle (oracle/port/ts_port.c, pinned to the reference's
bb_de_header) and the reference's own golden datagrams -- byte-exact, incl. dropped frames, resynchronisation, state
carried across calls, and the BBFRAMEs of a whole decoded T2 frame (config 4 chain)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.test_oracle_ts import CASES, GOLD, make, port_datagrams

pytestmark = pytest.mark.gpu

HEM_CASES = [c for c in CASES if CASES[c]['hem']]


@pytest.mark.parametrize('case', HEM_CASES)
def test_hem_datagrams_match_reference_golden_and_port(engine, case):
    frames, _ = make(case)
    g = np.load(GOLD)
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(frames)
    want = port_datagrams(frames)
    if case == 'hem_faults':
        # frame 2's flipped CRC bit turns its residue into the normal-mode value: the reference decodes it as normal mode,
        # and so does the general path the batch is switched to (status 3 = "normal-mode frame")
        assert st[2] == 3 and st[4] == 2
        kept = [d for d in want if d is not None]
        assert list(dl) == [0 if d is None else len(d) for d in want]
        assert list(g[case + '_len']) == [len(d) for d in kept]
        assert np.array_equal(ts, g[case + '_ts'])
        return
    assert (st == 0).all()
    assert list(dl) == [len(d) for d in want] == list(g[case + '_len'])
    assert np.array_equal(ts, g[case + '_ts'])


def test_faults_without_mode_flip(engine):
    """dropped frame (SYNCD 65535), too-long and too-short SYNCD after a split: resync branches, 0xF0 fill"""
    from tests.ts_helpers import bbframes
    frames, _ = bbframes(9552, 1180, 10, True, np.random.default_rng(3),
                         faults=((3, 'syncd65535'), (5, 'syncd_plus'), (7, 'syncd_minus')))
    frames[1, 79] ^= 1
    frames[1, 78] ^= 1                                   # header CRC error that is neither mode's residue
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(frames)
    want = port_datagrams(frames)
    assert st[1] == 1 and st[3] == 2 and want[1] is None and want[3] is None
    assert list(dl) == [0 if d is None else len(d) for d in want]
    assert np.array_equal(ts, np.concatenate([d for d in want if d is not None]))


def test_state_carries_across_calls_and_device_buffers(engine):
    import torch
    frames, packets = make('hem_normal_fec')
    engine.ts_reset(0)
    parts = []
    for a, b in ((0, 1), (1, 4), (4, 9)):
        ts, dl, st = engine.ts_packetize(torch.from_numpy(frames[a:b]).cuda())
        parts.append(ts.cpu().numpy())
    got = np.concatenate(parts)
    want = np.concatenate(port_datagrams(frames))
    assert np.array_equal(got, want)
    assert np.array_equal(got, packets.reshape(-1)[:len(got)])        # = the transmitted TS


def test_chain_bbframes_to_ts(engine):
    """config 4 geometry end to end: IQ -> ... -> BBFRAME bits on the GPU -> TS; equals the oracle's TS of the
    transmitted BBFRAMEs"""
    import torch
    from sdr_receiver_dvb_t2_b200.chain import FrameChain
    from tests.eq_helpers import tables
    from tools.modulator import Modulator
    t = tables('c16')
    m = Modulator(t, mod=2, cod=1, fec_normal=False, n_blocks=96, ti_len=3, seed=4)
    f = m.frame(noise_cn_db=15.0)
    ch = FrameChain(engine, t, mod=2, cod=1, fec_type=0, n_blocks=96, ti_len=3)
    r = ch.decode_frames(torch.from_numpy(f['time'][None]).cuda())
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(r['bits'])
    want = port_datagrams(f['bb'])
    assert (st == 0).all() and list(dl) == [len(d) for d in want]
    assert np.array_equal(ts.cpu().numpy(), np.concatenate(want))


def test_oversize_dfl_is_dropped_without_touching_the_state(engine):
    """a header with a valid CRC-8 that announces a data field longer than the frame (80 + DFL > k_bch): status 4, no
    datagram, and the frames around it come out as if it had been a header-CRC drop (the reference has no such check and
    would read past the frame)"""
    from tests.ts_helpers import bbframes, header
    frames, _ = bbframes(9552, 1180, 8, True, np.random.default_rng(5))
    bad = frames.copy()
    bad[3, :80] = header(65528, 0, True)                  # DFL = 65528 bits on a 9552-bit frame, CRC-8 correct
    bad[7, :80] = header(9552 - 80 + 8, 0, True)          # the last frame of the batch: one byte too long
    ref = frames.copy()
    for f in (3, 7):
        ref[f, 79] ^= 1
        ref[f, 78] ^= 1                                   # the same frames with a broken header CRC instead
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(bad)
    engine.ts_reset(0)
    ts2, dl2, st2 = engine.ts_packetize(ref)
    assert list(st) == [0, 0, 0, 4, 0, 0, 0, 4] and list(st2) == [0, 0, 0, 1, 0, 0, 0, 1]
    assert dl[3] == 0 and dl[7] == 0 and list(dl) == list(dl2) and np.array_equal(ts, ts2)
    want = port_datagrams(ref)
    assert np.array_equal(ts, np.concatenate([d for d in want if d is not None]))


NM_CASES = [c for c in CASES if not CASES[c]['hem']]


@pytest.mark.parametrize('case', NM_CASES)
def test_normal_mode_datagrams_match_reference_golden_and_port(engine, case):
    """normal mode (bb_de_header.cpp:166-331): CRC-8 of every packet checked against the byte on air, sync bytes re-inserted,
    the reference's reads behind the data field included -- byte for byte the golden datagrams of the reference"""
    frames, _ = make(case)
    g = np.load(GOLD)
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(frames)
    want = port_datagrams(frames)
    assert all(s in (2, 3) for s in st)
    assert list(dl) == [0 if d is None else len(d) for d in want]
    assert [n for n in dl if n] == list(g[case + '_len']) or list(g[case + '_len']) == [len(d) for d in want if d is not None]
    assert np.array_equal(ts, g[case + '_ts'])


@pytest.mark.parametrize('seed', range(12))
def test_mixed_mode_streams_equal_the_oracle(engine, seed):
    """HEM and NM stretches back to back with header faults of every kind and payload bit errors (transport_error_indicator
    set by the CRC check), in one call and cut into three calls (the packet state, the held-back tail and the CRC in flight
    carry over)"""
    from tests.test_ts_general_emu import emu_datagrams, mixed_stream
    frames = mixed_stream(seed)
    frames = frames[:len(emu_datagrams(frames))]          # up to where the reference's behaviour is defined (see there)
    want = port_datagrams(frames)
    flat = np.concatenate([d for d in want if d is not None] + [np.zeros(0, np.uint8)])
    engine.ts_reset(0)
    ts, dl, st = engine.ts_packetize(frames)
    assert list(dl) == [0 if d is None else len(d) for d in want]
    assert np.array_equal(ts, flat)
    engine.ts_reset(0)
    parts = []
    n = len(frames)
    for a, b in ((0, n // 3), (n // 3, n // 3 + 1), (n // 3 + 1, n)):
        if b > a:
            parts.append(engine.ts_packetize(frames[a:b])[0])
    assert np.array_equal(np.concatenate(parts), flat)


def test_parallel_path_continues_from_the_state_the_general_path_leaves(engine):
    """'hem_faults' holds one frame a header bit error turned into normal mode: the call with that frame runs on the general
    path and leaves a regular packet state, the following pure-HEM calls run the parallel scan again from its held-back tail
    (and a call of the general path continues from a tail the parallel path left)"""
    frames, _ = make('hem_faults')
    g = np.load(GOLD)
    engine.ts_reset(0)
    parts = [engine.ts_packetize(frames[a:b])[0] for a, b in ((0, 2), (2, 3), (3, 6), (6, 10))]
    assert np.array_equal(np.concatenate(parts), g['hem_faults_ts'])
