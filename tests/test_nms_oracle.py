"""CPU: oracle/nms_oracle.c (plain-C twin) and the numpy twin against the compiled, unmodified
reference extension oracle/_ref/nms_1d_cpu_vg.so (built by oracle/build_ref.py from
/root/reference/libs/nms/src/nms_cpu.cpp).  Skipped where the binary is absent."""
import numpy as np
import pytest

from oracle import build_ref, nms_oracle
from oracle import grounder_oracle as go


def _cands(rng, n, T=2304.0, quant=None):
    c = rng.uniform(0, T, n).astype(np.float32)
    l = rng.uniform(1, 200, n).astype(np.float32)
    segs = np.stack([c - l / 2, c + l / 2], 1).astype(np.float32)
    sc = rng.uniform(0, 1, n).astype(np.float32)
    if quant:
        sc = (np.round(sc * quant) / quant).astype(np.float32)
    return segs, sc


@pytest.fixture(scope='module')
def ref():
    build_ref.build_reference_nms()
    s, n = nms_oracle.reference_fns()
    if s is None:
        pytest.skip('compiled reference extension not available')
    return s, n


@pytest.mark.parametrize('n,sigma,min_score,method', [
    (1, 0.9, 1e-3, 2), (2, 0.9, 1e-3, 2), (57, 0.9, 1e-3, 2), (300, 0.5, 0.05, 2),
    (300, 0.9, 0.3, 2), (500, 0.9, 1e-3, 1), (500, 0.9, 1e-3, 0), (2000, 0.9, 1e-3, 2)])
def test_softnms_c_twin_matches_reference(ref, n, sigma, min_score, method):
    rng = np.random.default_rng(n * 7 + method)
    segs, sc = _cands(rng, n, T=400.0 if n <= 500 else 2304.0)
    d_ref, i_ref = ref[0](segs, sc, 0.1, sigma, min_score, method)
    d_c, i_c = nms_oracle.softnms(segs, sc, 0.1, sigma, min_score, method)
    assert np.array_equal(i_ref, i_c)
    assert np.array_equal(d_ref, d_c)           # same libm expf, same op order -> bit-exact
    # truncated run == prefix of the full run
    d5, i5 = nms_oracle.softnms(segs, sc, 0.1, sigma, min_score, method, max_iters=5)
    k = min(5, len(i_ref))
    assert np.array_equal(d5, d_ref[:k]) and np.array_equal(i5, i_ref[:k])


def test_softnms_numpy_twin_small(ref):
    rng = np.random.default_rng(3)
    for n in (1, 5, 40, 120):
        segs, sc = _cands(rng, n, T=200.0, quant=20)       # heavy ties
        d_ref, i_ref = ref[0](segs, sc, 0.1, 0.9, 0.2, 2)
        d_np, i_np = go.soft_nms_np(segs, sc, 0.1, 0.9, 0.2, 2)
        assert np.array_equal(i_ref, i_np)
        np.testing.assert_allclose(d_ref, d_np, rtol=2e-6, atol=0)   # numpy exp vs glibc expf: ulp-level per decay
        d_c, i_c = nms_oracle.softnms(segs, sc, 0.1, 0.9, 0.2, 2)
        assert np.array_equal(i_ref, i_c) and np.array_equal(d_ref, d_c)


@pytest.mark.parametrize('n', [1, 33, 500, 2000])
def test_hardnms_twins_match_reference(ref, n):
    rng = np.random.default_rng(n)
    segs, sc = _cands(rng, n)                                  # tie-free
    k_ref = ref[1](segs, sc, 0.5)
    assert np.array_equal(k_ref, nms_oracle.nms(segs, sc, 0.5))
    if n <= 500:
        assert np.array_equal(k_ref, go.hard_nms_np(segs, sc, 0.5))


def test_empty():
    d, i = nms_oracle.softnms(np.zeros((0, 2), np.float32), np.zeros(0, np.float32), 0.1, 0.9, 1e-3, 2)
    assert len(d) == 0 and len(i) == 0
    assert len(nms_oracle.nms(np.zeros((0, 2), np.float32), np.zeros(0, np.float32), 0.5)) == 0
