"""The exact parallel evaluation of the reference's ordered float sums (demap_sum_chunks_kernel / demap_sum_stitch_kernel, K4) follows a bit-level
model (tools/ordered_sum_model.py): inside a binade the running sum is an integer in ulps, round-to-nearest-even addition
of a non-negative term depends on the sum only through its parity, runs of terms compose as (delta-if-even, delta-if-odd)
pairs, binade crossings are done as real float additions.  This pins the model against a serial float32 sum on adversarial
inputs (exact ties, tiny and huge terms, zeros); the GPU tests pin the kernel against the oracle."""
import numpy as np

from tools import ordered_sum_model as M


def test_model_equals_serial_float32_sum():
    assert M.check(trials=120, seed=3) == 120


def test_model_chunkings_agree():
    rng = np.random.default_rng(9)
    t = M.random_terms(1, 2500, rng)                     # many exact ties
    want = M.fbits(M.serial(t))
    for ch, e in ((64, 4), (256, 8), (1024, 16), (8, 1)):
        got, _ = M.ordered_sum(t, CH=ch, E=e)
        assert M.fbits(got) == want


def test_stitched_scheme_takes_most_chunks_as_one_integer():
    """demapper-like terms (squares of small numbers): the prediction holds for nearly every chunk, the result is exact"""
    rng = np.random.default_rng(4)
    t = ((rng.standard_normal(40000) * 0.05) ** 2 + 0.3).astype(np.float32)
    got, fast, slow = M.stitched_sum(t, chunk=256)
    assert M.fbits(got) == M.fbits(M.serial(t))
    assert fast > 5 * slow
