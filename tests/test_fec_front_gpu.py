"""K3 / K4 parity through the C-ABI: time de-interleaver + Q-delay removal (exact: a permutation) and the
soft demapper.  The demapper has ONE order-dependent float reduction (sum_s / sum_e -> precision); the CUDA path
accumulates it in the reference's order (float, cell by cell), so precision and every int8 LLR must be identical
to the oracle -- both with the oracle's precision handed in and with the GPU's own."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.fec_helpers import CONFIGS, config_input, cpf_of, port_chain, sha

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name', ['A_s64_r35', 'D_sqpsk_r12', 'G_s64_r23', 'B_n16_r12', 'F_n64_r35', 'H_n256_r34',
                                  'E_n256_r23'])
def test_ti_and_demap_match_oracle(engine, name):
    cfg = CONFIGS[name]
    stream, blocks = config_input(name)
    ti_ref, llr_ref, snr_ref, prec_ref = port_chain(name, stream, blocks)
    engine.ti_configure(0, cfg['fec'], cfg['mod'], max(blocks))
    ti = engine.ti_deinterleave(0, stream, blocks)
    assert np.array_equal(ti.view(np.float32), ti_ref.view(np.float32))
    # (1) pinned precision -> bit-exact LLRs and in-place derotation
    cells = ti.copy()
    r = engine.demap(cells, blocks, cfg['mod'], cfg['rot'], cfg['fec'], cfg['cod'], precision_in=prec_ref)
    assert np.array_equal(r['llr'], llr_ref)
    # (2) own precision
    cells2 = ti.copy()
    r2 = engine.demap(cells2, blocks, cfg['mod'], cfg['rot'], cfg['fec'], cfg['cod'])
    assert np.array_equal(r2['precision'], prec_ref)
    assert np.allclose(r2['snr'], snr_ref, rtol=0, atol=1e-4)        # log10f: library ulp
    assert np.array_equal(r2['llr'], llr_ref)
    if cfg['rot']:
        # the reference derotates its input buffer in place (llr_demapper.cpp:555-557)
        off = 0
        _, _, _, derot = O.port_demap(ti_ref[:blocks[0] * cpf_of(cfg)], cfg['mod'], cfg['rot'], cfg['fec'], cfg['cod'])
        assert np.array_equal(cells[:len(derot)].view(np.float32), derot.view(np.float32))


def test_golden_digests_through_gpu(engine):
    """the reference's own digests, straight against the CUDA path (pinned precision)"""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'fec_ref.npz'))
    for name in ('A_s64_r35', 'F_n64_r35'):
        cfg = CONFIGS[name]
        stream, blocks = config_input(name)
        if sha(stream) != str(g[name + '_in_sha']):
            pytest.skip('numpy generator differs')
        _, _, _, prec = port_chain(name, stream, blocks)
        engine.ti_configure(1, cfg['fec'], cfg['mod'], max(blocks))
        ti = engine.ti_deinterleave(1, stream, blocks)
        assert sha(ti) == str(g[name + '_ti_sha'])
        r = engine.demap(ti, blocks, cfg['mod'], cfg['rot'], cfg['fec'], cfg['cod'], precision_in=prec)
        assert sha(r['llr'].reshape(-1)[:int(g[name + '_n_llr'])]) == str(g[name + '_llr_sha'])


def test_short_frames_16qam_256qam_and_device_buffers(engine):
    """geometries the reference cannot run (its 4-cell loop never sees the end of a 16 200-bit frame for
    16/256-QAM): checked against the oracle restatement, which has no such limit; device-resident I/O."""
    import torch
    rng = np.random.default_rng(3)
    for mod, cod in ((1, 0), (3, 2)):
        cpf = 16200 // (2 * (mod + 1))
        blocks = [3, 2]
        n = sum(blocks) * cpf
        stream = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * 0.7
        perm = O.port_cell_permutation(3, cpf)
        ti_ref = O.port_ti_blocks(stream, blocks, cpf, perm, [0, 0.0])
        engine.ti_configure(2, 0, mod, 3)
        d = torch.from_numpy(stream).cuda()
        ti = engine.ti_deinterleave(2, d, blocks)
        engine.sync()
        assert np.array_equal(ti.cpu().numpy().view(np.float32), ti_ref.view(np.float32))
        off = 0
        precs, llrs = [], []
        for nf in blocks:
            llr, snr, p, _ = O.port_demap(ti_ref[off:off + nf * cpf], mod, 1, 0, cod)
            precs.append(p), llrs.append(llr)
            off += nf * cpf
        r = engine.demap(ti, blocks, mod, 1, 0, cod, precision_in=np.array(precs, np.float32))
        engine.sync()
        assert np.array_equal(r['llr'].cpu().numpy(), np.concatenate(llrs))


def test_ti_edge_cases(engine):
    engine.ti_configure(3, 0, 2, 4)
    with pytest.raises(Exception):
        engine.ti_deinterleave(3, np.zeros(5 * 2700, np.complex64), [5])      # more blocks than configured
    with pytest.raises(Exception):
        engine.ti_deinterleave(9, np.zeros(2700, np.complex64), [1])          # PLP never configured
    out = engine.ti_deinterleave(3, np.zeros(0, np.complex64), [])
    assert out.shape == (0,)
