"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/kmdiff_b200.h declares,
refuses to compute without a device (no CPU fallback), and its host-only helpers (table, chi-square,
corrector predicates, k-mer / FASTA formatting, synthetic rows) agree with the oracle.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from kmdiff_b200 import _lib
from conftest import SHAPES, make_cohort

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "kmdiff_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # and the Python binding knows all of them
    assert sorted(_lib.SYMBOLS) == syms
    assert L.kmd_abi_version() == _lib.KMD_ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Config) == 80
    assert C.sizeof(_lib.Counters) == 40
    assert C.sizeof(_lib.Synth) == 40
    assert _lib.SURVIVOR_DTYPE.itemsize == 56


@pytest.mark.skipif(_lib.load().kmd_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(K.KmdError) as e:
        K.DiffContext(10, 10, 10**10, 10**10)
    assert e.value.code == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_create_argument_checks():
    L = _lib.load()
    h = C.c_void_p()
    bad = _lib.Config(_lib.KMD_ABI_VERSION + 1, 0, 1, 1, 31, 1, 4, 0, 1, 1, 1e-6, None, 0, 0)
    assert L.kmd_create(C.byref(bad), C.byref(h)) == _lib.ERR_ARG
    for field, value in (("nb_controls", 0), ("kmer_words", 3), ("count_bytes", 3)):
        cfg = _lib.Config(_lib.KMD_ABI_VERSION, 0, 1, 1, 31, 1, 4, 0, 1, 1, 1e-6, None, 0, 0)
        setattr(cfg, field, value)
        assert L.kmd_create(C.byref(cfg), C.byref(h)) == _lib.ERR_ARG
        assert L.kmd_last_error(None)


def test_lf_table_is_bit_identical_to_oracle():
    t = K.LogFactorialTable(12000)
    assert np.array_equal(t.table, oracle.lf_table(12000))
    assert t[10] == 15.104412573075514 and t[50] == 148.47776695177302  # reference tests/factorial_test.cpp


def test_chisqc_host_matches_oracle():
    L = _lib.load()
    xs = np.exp(np.linspace(np.log(1e-8), np.log(1500.0), 3000))
    got = np.array([L.kmd_chisqc(1.0, float(x)) for x in xs])
    want = np.array([oracle.chisqc(1.0, float(x)) for x in xs])
    assert np.array_equal(got, want)
    assert L.kmd_chisqc(1.0, 0.0) == 1.0 and L.kmd_chisqc(1.0, -1.0) == 1.0


@pytest.mark.parametrize("kind", ["nothing", "bonferroni", "benjamini", "sidak", "holm"])
def test_corrector_mirrors(kind):
    """reference tests/corrector_test.cpp semantics through the Python mirror (stateful like the reference)"""
    rng = np.random.default_rng(3)
    ps = np.sort(np.exp(rng.uniform(np.log(1e-14), np.log(1e-3), 400)))
    for total in (300, 10**6, 10**9):
        thr = 1e-6 if kind == "nothing" else 0.05
        ours = K.make_corrector(K.correction_from_str("disabled" if kind == "nothing" else kind), thr, total)
        want = oracle.Corrector(oracle.CORRECTIONS[kind], thr, total)
        assert [ours.apply(float(p)) for p in ps] == [want.apply(float(p)) for p in ps]


def test_corrector_reference_goldens():
    assert K.basic_threshold(0.05).apply(0.04) and not K.basic_threshold(0.05).apply(0.06)
    assert K.bonferroni(0.05, 100).apply(0.0004) and not K.bonferroni(0.05, 100).apply(0.0006)
    c = K.benjamini(0.25, 25)
    assert c.apply(0.009) and not c.apply(0.02)
    assert K.sidak(0.05, 100).apply(0.00050) and not K.sidak(0.05, 100).apply(0.00052)
    h = K.holm(0.05, 100)
    for _ in range(90):
        h.apply(0)
    assert h.apply(0.004) and not h.apply(0.006)


def test_first_reject_equals_oracle_aggregate():
    rng = np.random.default_rng(8)
    p = np.sort(np.exp(rng.uniform(np.log(1e-13), np.log(1e-5), 1000)))
    for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
        for total in (1000, 10**7, 10**10):
            n, cut = K.first_reject(p, kind, 0.05, total)
            keep = oracle.aggregate(p, int(kind), 0.05, total)
            assert n == len(keep)
            assert cut == (p[n] if n < len(p) else np.inf)


def test_kmer_to_string_matches_oracle():
    rng = np.random.default_rng(1)
    for k in (1, 20, 31, 32, 33, 63, 64):
        kw = (k + 31) // 32
        for _ in range(20):
            w = rng.integers(0, 2**63, size=kw, dtype=np.uint64)
            assert K.kmer_to_string(w, k) == oracle.kmer_to_string(w, k)


def test_format_fasta_matches_reference_format():
    rec = np.zeros(1, dtype=_lib.SURVIVOR_DTYPE)[0]
    cases = [(1.11209e-10, 0.0, 30.0), (1.0429e-17, 52.9, 37.2), (0.0, 1e15, 123456789.0), (4.9e-324, 3.99, 0.1),
             (1.7928739662752438e-42, 37.2, 280.0), (5e-7, 0.5, 1e16), (1e-100, 2.5, 1e-5)]
    for k in (20, 31, 63):
        for i, (p, mc, kk) in enumerate(cases):
            rec["kmer"] = [0x1B1B1B1B1B1B1B1B, 0x0123456789ABCDEF]
            rec["p_value"], rec["mean_control"], rec["mean_case"] = p, mc, kk
            seq = oracle.kmer_to_string(rec["kmer"][: (k + 31) // 32], k)
            assert K.format_fasta(i, rec, k) == oracle.fasta_record(i, p, mc, kk, seq)


def test_synthetic_rows_are_valid_kmtricks_rows():
    for shape in SHAPES:
        nc, nk, k, b = SHAPES[shape]
        spec, rows, totals = make_cohort(shape, 5000, seed=3)
        R = spec.row_bytes
        m = rows.reshape(5000, R)
        kw = spec.kmer_words
        km = m[:, :8 * kw].copy().view(np.uint64).reshape(5000, kw)
        key = km[:, kw - 1].astype(object) * (1 << 64 if kw == 2 else 1) + (km[:, 0].astype(object) if kw == 2 else 0)
        assert all(key[i] < key[i + 1] for i in range(4999)), "rows of a partition are sorted by k-mer"
        if 2 * k < 64 * kw:
            assert int(km[:, kw - 1].max()) < (1 << (2 * k - 64 * (kw - 1)))
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[b]
        counts = m[:, 8 * kw:].copy().view(dt).reshape(5000, nc + nk)
        assert (counts.sum(axis=1) > 0).all(), "kmtricks never emits all-zero rows"
        # deterministic & sliceable
        again = spec.rows_host(100, 1234, threads=1)
        assert np.array_equal(again, spec.rows_host(5000, 0, threads=3).reshape(5000, R)[1234:1334].ravel())
        assert len(totals) == nc + nk


def test_oracle_sees_enrichment_in_synthetic_rows():
    spec, rows, totals = make_cohort("c2", 100_000, seed=1, enriched_ppm=1000)
    want, n_ctl, n_case = oracle.Model(10, 10, totals).diff(rows, 1, 4, 0.05 / 100000)
    assert 50 <= len(want) <= 200 and n_case > n_ctl > 0
