"""K1 parity through the C-ABI: batched CUDA FFT vs the float64 DFT oracle.  FFTW (the reference's FFT) is a
binary-only dependency, so the contract is the stated tolerance: <= 1e-5 * max|X|; observed ~3e-7."""
import numpy as np
import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('n', [32768, 16384, 8192, 4096, 2048, 1024, 256])      # 1024: the P1 symbol (p1_symbol.cpp:34-35)
def test_fft_matches_float64_dft(engine, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    x[1] = 0
    x[1, 5] = 1.0                                    # impulse: every bin has magnitude 1
    x[2] = np.exp(2j * np.pi * 37 * np.arange(n) / n).astype(np.complex64)   # one carrier -> one bin
    got = engine.fft(x)
    for b in range(3):
        want = O.port_fft(x[b]).astype(np.complex128)
        assert np.abs(got[b] - want).max() <= TOL * np.abs(want).max()
    assert np.argmax(np.abs(got[2])) == n // 2 + 37                 # carrier 37 right of DC after the half swap
    assert abs(np.abs(got[2]).max() - n) < 1e-2 * n ** 0.5          # unnormalised


def test_fft_batch_spanning_chunks_and_device_buffers(engine):
    """batch larger than one L2 chunk, data resident on the device; size-independent properties: every symbol equals
    the oracle on a sample, Parseval holds for all"""
    import torch
    n, batch = 32768, 130
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    x = torch.randn((batch, n, 2), generator=g, device='cuda')
    xc = torch.view_as_complex(x).contiguous()
    y = engine.fft(xc)
    engine.sync()
    e_in = (xc.abs() ** 2).sum(dim=1).double()
    e_out = (y.abs() ** 2).sum(dim=1).double()
    assert torch.allclose(e_out, e_in * n, rtol=1e-5)
    for b in (0, 95, 96, 129):
        want = O.port_fft(xc[b].cpu().numpy())
        got = y[b].cpu().numpy()
        assert np.abs(got - want).max() <= TOL * np.abs(want).max()


def test_fft_rejects_bad_sizes(engine):
    with pytest.raises(Exception):
        engine.fft(np.zeros((1, 1000), np.complex64))
    assert engine.fft(np.zeros((0, 4096), np.complex64)).shape == (0, 4096)
