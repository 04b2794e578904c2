"""Shared helpers for the TI / demapper parity tests: regenerate the seeded inputs of
tools/make_golden_fec.py and run the CPU oracle (port) over them."""
import hashlib

import numpy as np

from oracle import pyoracle as O
from tools.make_golden_fec import CONFIGS, frame_streams, ti_split


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cpf_of(cfg):
    return (64800 if cfg['fec'] else 16200) // (2 * (cfg['mod'] + 1))


def config_input(name):
    """-> (stream complex64 of all frames' PLP cells in arrival order, n_fec per TI block list)"""
    cfg = CONFIGS[name]
    streams = []
    for p2, syms in frame_streams(name, cfg):
        streams.append(np.concatenate([p2[2200:]] + syms))
    return np.concatenate(streams), ti_split(cfg['nb'], cfg['ti_len']) * cfg['frames']


def port_chain(name, stream, blocks):
    """port oracle: TI de-interleave then demap every TI block -> (ti cells, llr [n_fec][N], snr[], precision[])"""
    cfg = CONFIGS[name]
    cpf = cpf_of(cfg)
    perm = O.port_cell_permutation(max(blocks), cpf)
    ti = O.port_ti_blocks(stream, blocks, cpf, perm, [0, 0.0])
    llrs, snrs, precs, off = [], [], [], 0
    for nf in blocks:
        llr, snr, p, _ = O.port_demap(ti[off:off + nf * cpf], cfg['mod'], cfg['rot'], cfg['fec'], cfg['cod'])
        llrs.append(llr), snrs.append(snr), precs.append(p)
        off += nf * cpf
    return ti, np.concatenate(llrs), np.array(snrs, np.float32), np.array(precs, np.float32)
