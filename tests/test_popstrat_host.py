"""Host-side pieces of the --pop-correction path that live in the product library (no GPU needed):
the jump-ahead Sampler, the symmetric eigensolver and the smartpca PC read-back, each against the
oracle (kmd_popstrat_oracle.c) and numpy."""
import json
import os

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "popstrat_vectors.json")) as f:
    GOLD = json.load(f)


@pytest.mark.parametrize("case", GOLD["sampler"], ids=lambda c: f"seed{c['seed']}")
def test_sampler_golden(case):
    """golden draws of the unmodified reference Sampler (libstdc++ minstd_rand0 + generate_canonical)"""
    got = np.flatnonzero(K.sampler_rows(case["seed"], case["kmer_pca"], 0, case["n"]))
    assert got.tolist() == case["taken"]


@pytest.mark.parametrize("seed,v", [(0, 0.03), (7, 0.5), (2**31 - 1, 0.001), (2**31 - 2, 0.2), (2**40 + 17, 0.1),
                                    (42, 0.0), (42, 1.0), (42, 1e-9)])
def test_sampler_jump_ahead_matches_oracle(seed, v):
    n = 40000
    full = oracle.sampler_rows(seed, v, n)
    assert np.array_equal(K.sampler_rows(seed, v, 0, n).astype(bool), full)
    # the engine has period 2^31 - 2 and a row takes two calls: decisions repeat every 2^30 - 1 rows
    period = 2**30 - 1
    for first in (1, 777, 39999, period - 100, 2**31 - 2, 2**31 - 1, 2**33 + 5, 2**40 + 12345):
        part = K.sampler_rows(seed, v, first, 300).astype(bool)
        f = first % period
        if f + 300 <= n:
            assert np.array_equal(part, full[f:f + 300])
        elif f < n:
            assert np.array_equal(part[:n - f], full[f:])
        else:   # wraps around the period: its tail is the start of the sequence
            k = period - f
            if k < 300:
                assert np.array_equal(part[k:], full[:300 - k])


def test_sym_eig_matches_numpy_and_oracle():
    rng = np.random.default_rng(9)
    for n in (1, 2, 7, 40, 130):
        x = rng.standard_normal((n, 2 * n + 3))
        a = x @ x.T / (2 * n)
        ev, vec = K.sym_eig(a)
        w = np.linalg.eigvalsh(a)[::-1]
        assert np.allclose(ev, w, rtol=1e-12, atol=1e-13)
        assert np.abs(a @ vec.T - vec.T * ev).max() < 1e-12 * max(1.0, w[0])
        assert np.abs(vec @ vec.T - np.eye(n)).max() < 1e-12
        oe, ov = oracle.sym_eig(a)
        assert np.allclose(ev, oe, rtol=1e-11, atol=1e-13)
        for j in range(min(n, 10)):
            assert abs(vec[j] @ ov[j] - 1) < 1e-9   # same sign convention


def _geno(rng, m, n):
    f = rng.beta(0.5, 0.5, size=(m, 1))
    g = (rng.random((m, n)) < f).astype(np.uint8)
    g[0] = 0
    g[1] = 1
    g[:, :n // 3] |= (rng.random((m, 1)) < 0.15).astype(np.uint8)   # a sub-population
    return g


def test_smartpca_pcs_matches_oracle():
    rng = np.random.default_rng(21)
    m, n = 5000, 36
    g = _geno(rng, m, n)
    xtx, _ = oracle.smartpca_xtx(g)
    oe, ov = oracle.sym_eig(xtx)
    want = oracle.smartpca_pcs(g, ov, 10)
    ev, pcs = K.smartpca_pcs(xtx, 10)
    assert np.allclose(ev, oe, rtol=1e-10, atol=1e-12)
    assert pcs.shape == (n, 10)
    assert np.array_equal(pcs, np.round(pcs, 4))
    # the same "%10.4f" read-back; a coordinate may differ by one quantum where it sits on a rounding edge
    assert np.abs(pcs - want).max() <= 1e-4 + 1e-12
    assert np.mean(pcs != want) < 0.01


def test_smartpca_normalize():
    rng = np.random.default_rng(2)
    a = rng.random((9, 9))
    a = a + a.T
    b = a.copy()
    from kmdiff_b200 import _lib
    _lib.load().kmd_smartpca_normalize(b.ctypes.data, 9)
    assert np.allclose(b, a / (np.trace(a) / 8), rtol=1e-15)
