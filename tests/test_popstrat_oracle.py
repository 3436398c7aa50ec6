"""Pins oracle/kmd_popstrat_oracle.c (sampler, IRLS, pop_strat_corrector) against golden vectors made
from the unmodified reference (tests/golden/popstrat_vectors.json <- make_popstrat_golden.py) and,
where oracle/_ref exists, against the live reference library; checks the smartpca restatement against
numpy (it is pinned against the real smartpca in tests/test_smartpca_reference.py).  CPU only."""
import json
import os
import tempfile

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "popstrat_vectors.json")) as f:
    GOLD = json.load(f)

needs_ref = pytest.mark.skipif(not (oracle.ref_available() and hasattr(oracle.ref(), "ref_glm_irls")),
                               reason="oracle/_ref not built (no /root/reference here)")


def unhex(v, shape=None):
    a = np.array([float.fromhex(x) for x in v], dtype=np.float64)
    return a if shape is None else a.reshape(shape)


@pytest.mark.parametrize("case", GOLD["sampler"], ids=lambda c: f"seed{c['seed']}")
def test_sampler_golden(case):
    got = np.flatnonzero(oracle.sampler_rows(case["seed"], case["kmer_pca"], case["n"]))
    assert got.tolist() == case["taken"]


def test_sampler_skip_ahead_matches_sequential():
    full = oracle.sampler_rows(42, 0.2, 20000)
    for first in (0, 1, 777, 19999):
        part = oracle.sampler_rows(42, 0.2, min(500, 20000 - first), first_row=first)
        assert np.array_equal(part, full[first:first + part.size])


@pytest.mark.parametrize("case", GOLD["irls"], ids=lambda c: c["kind"] + str(c["p"]))
def test_irls_golden(case):
    x = unhex(case["x"], (case["n"], case["p"]))
    y = unhex(case["y"])
    w, s, q, e, it = oracle.glm_irls(x, y)
    assert (s, q, it) == (case["singular"], case["nan"], case["iters"])
    assert np.array_equal(w, unhex(case["weights"]), equal_nan=True)
    assert e == float.fromhex(case["error"]) or (np.isnan(e) and np.isnan(float.fromhex(case["error"])))


@pytest.mark.parametrize("case", GOLD["inverse"], ids=lambda c: f"n{c['n']}")
def test_inverse_golden(case):
    n = case["n"]
    inv, s, q = oracle.inverse(unhex(case["a"], (n, n)))
    assert (s, q) == (case["singular"], case["nan"])
    assert np.array_equal(inv, unhex(case["inv"], (n, n)))


@pytest.mark.parametrize("case", GOLD["popstrat"], ids=lambda c: f"{c['nb_controls']}+{c['nb_cases']}pc{c['npc']}")
def test_popstrat_golden(case):
    nc, nk, npc = case["nb_controls"], case["nb_cases"], case["npc"]
    n = nc + nk
    ps = oracle.PopStrat(nc, nk, case["totals"], unhex(case["Z"], (n, 10)), npc)
    assert np.array_equal(ps.null_features, unhex(case["null_features"], (n, npc + 2)))
    assert np.array_equal(ps.alt_features, unhex(case["alt_features"], (n, npc + 3)))
    assert np.array_equal(ps.null_model, unhex(case["null_model"]))
    counts = unhex(case["counts"], (-1, n))
    want = unhex(case["p"])
    got = np.array([ps.apply(c) for c in counts])
    assert np.array_equal(got, want, equal_nan=True)
    if n >= 50:   # (a 24-sample cohort with 12 features is separable by the null model: every p ~ 1)
        assert np.any(want < 1e-3) and np.any(want > 1e-2)


@needs_ref
def test_sampler_live_reference():
    for seed, v in ((0, 0.03), (7, 0.5), (2**31 - 1, 0.001), (2**40 + 17, 0.1)):
        n = 50000
        out = np.zeros(n, np.uint8)
        oracle.ref().ref_sampler_rows(seed, v, n, out.ctypes.data)
        assert np.array_equal(oracle.sampler_rows(seed, v, n), out.astype(bool))


@needs_ref
def test_irls_live_reference():
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(8, 120))
        p = int(rng.integers(2, 14))
        x = np.c_[np.ones(n), rng.normal(size=(n, p - 1)) * 10.0 ** rng.integers(-9, 10, size=p - 1)]
        y = (rng.random(n) < 0.5).astype(np.float64)
        if trial % 5 == 0:
            y = (x[:, 1] > 0).astype(np.float64)         # separable: mu saturates, rows get dropped
        mi = int(rng.integers(1, 101))
        a = oracle.glm_irls(x, y, max_iters=mi)
        b = oracle.glm_irls(x, y, max_iters=mi, impl="ref")
        assert a[1:3] == b[1:3] and a[4] == b[4]
        assert np.array_equal(a[0], b[0], equal_nan=True)


@needs_ref
def test_popstrat_live_reference():
    rng = np.random.default_rng(11)
    nc, nk, npc = 25, 35, 10
    n = nc + nk
    totals = rng.integers(int(5e8), int(2e9), n).astype(np.uint64)
    Z = np.round(rng.normal(scale=0.1, size=(n, 10)), 4)
    ps = oracle.PopStrat(nc, nk, totals, Z, npc)
    with tempfile.TemporaryDirectory() as d:
        rp = oracle.RefPopStrat(nc, nk, totals, Z, npc, d)
        assert np.array_equal(ps.null_features, rp.null_features)
        assert np.array_equal(ps.alt_features, rp.alt_features)
        assert np.array_equal(ps.null_model, rp.null_model)
        for t in range(60):
            c = rng.poisson(rng.gamma(2, 3), n).astype(np.float64)
            if t % 3 == 0:
                c[nc:] += rng.poisson(30, nk)
            if t % 7 == 0:
                c[:] = 0
            a, b = ps.apply(c), rp.apply(c)
            assert a == b or (np.isnan(a) and np.isnan(b))


def _geno(rng, m, n):
    f = rng.beta(0.5, 0.5, size=(m, 1))
    g = (rng.random((m, n)) < f).astype(np.uint8)
    g[0] = 0          # no carrier: dropped (smartpca.c:921)
    g[1] = 1          # every sample carries it: zero column after centring
    return g


def test_smartpca_xtx_matches_formula():
    rng = np.random.default_rng(3)
    m, n = 3000, 24
    g = _geno(rng, m, n)
    xtx, used = oracle.smartpca_xtx(g)
    c = g.sum(1)
    keep = c > 0
    assert used == int(keep.sum())
    mu = c[keep] / n
    pp = 1 - np.sqrt(1 - mu)
    y = pp * (1 - pp)
    f = np.where(y > 0, 1 / np.sqrt(np.where(y > 0, y, 1)), 1.0)
    X = (g[keep] - mu[:, None]) * f[:, None]
    want = X.T @ X
    want /= np.trace(want) / (n - 1)
    assert np.allclose(xtx, want, rtol=1e-11, atol=1e-13)
    assert np.array_equal(xtx, xtx.T)


def test_sym_eig_and_pcs():
    rng = np.random.default_rng(4)
    m, n = 4000, 30
    g = _geno(rng, m, n)
    g[:, :12] |= (rng.random((m, 1)) < 0.15).astype(np.uint8)     # structure: a sub-population
    xtx, _ = oracle.smartpca_xtx(g)
    evals, evecs = oracle.sym_eig(xtx)
    w, v = np.linalg.eigh(xtx)
    assert np.allclose(evals, w[::-1], rtol=1e-10, atol=1e-12)
    for j in range(10):
        assert abs(abs(evecs[j] @ v[:, n - 1 - j]) - 1) < 1e-9
        assert np.allclose(xtx @ evecs[j], evals[j] * evecs[j], atol=1e-10)
    pcs = oracle.smartpca_pcs(g, evecs, 10)
    # the printed coordinates are the unit eigenvectors quantised to 1e-4 (SURVEY.md §7 "GRM")
    assert np.abs(pcs - evecs[:10].T).max() <= 0.5e-4 + 1e-9
    assert np.array_equal(pcs, np.round(pcs, 4))


def test_reference_linear_test_goldens():
    """the golden values of the reference's own tests/linear_test.cpp (TEST(linear, utils) and
    TEST(linear, lu_decomposition)): sigmoid(1.0), predict({1,2,3},{1,2,3}), and the 4x4 inverse"""
    L = oracle.lib()
    assert abs(L.kmo_sigmoid(1.0) - 0.7310585786300048792512) < 1e-15
    assert abs(L.kmo_sigmoid(14.0) - 0.9999991684719723358679) < 1e-15       # predict = sigmoid(linear_predictor) = sigmoid(14)
    m = np.array([[1, 2, 1, 1], [1, 1, 6, 1], [1, 0, 1, 0], [1, 0, 1, 1]], dtype=np.float64)
    want = np.array([[0.1, -0.2, 1, 0.1], [0.5, 0, 0, -0.5], [-0.1, 0.2, 0, -0.1], [0, 0, -1, 1]])
    inv, singular, is_nan = oracle.inverse(m)
    assert not singular and not is_nan
    assert np.abs(inv - want).max() < 1e-15
    if oracle.ref_available() and hasattr(oracle.ref(), "ref_inverse"):
        rinv, _, _ = oracle.inverse(m, impl="ref")
        assert np.array_equal(inv, rinv)
