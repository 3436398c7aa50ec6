#!/usr/bin/env python
"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libkmdiff_ref.so, built
from /root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the committed JSON files are what travels.  Doubles are stored
as C99 hex-float strings so the fixtures are bit-exact.

Files:
  reference_tests.json  golden values the reference's own unit tests hold for this path
                        (tests/factorial_test.cpp, tests/model_test.cpp, tests/corrector_test.cpp)
                        plus the two known answers of SURVEY.md §8c — hand-transcribed, then
                        re-checked against the library below before being written.
  ref_vectors.json      outputs of the reference library on seeded inputs: log-factorial look-ups,
                        chi-square tail on a grid, PoissonLikelihood::process on small cohorts of
                        every count width, corrector predicates.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def cohort_rows(rng, nc, nk, b, n, kmer_words=1):
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[b]
    cmax = int(np.iinfo(dt).max)
    lam = rng.gamma(2.0, 3.0, size=(n, 1))
    counts = rng.poisson(lam, size=(n, nc + nk)).astype(np.int64)
    boost = rng.random(n) < 0.3
    counts[boost, nc:] += rng.poisson(25.0, size=(int(boost.sum()), nk))
    flip = rng.random(n) < 0.1
    counts[flip, :nc] += rng.poisson(25.0, size=(int(flip.sum()), nc))
    big = rng.random(n) < 0.05
    counts[big] *= int(min(cmax // 64, 3000))
    counts = np.minimum(counts, cmax)
    counts[0] = 0                      # an all-zero row: lambda <= 0 branches (model.hpp:135)
    counts[1, :nc] = 0                 # one empty cohort
    counts[2, nc:] = 0
    counts = counts.astype(dt)
    rows = np.zeros((n, 8 * kmer_words + (nc + nk) * b), dtype=np.uint8)
    km = rng.integers(0, 2**62, size=(n, kmer_words), dtype=np.uint64)
    rows[:, :8 * kmer_words] = km.view(np.uint8).reshape(n, -1)
    rows[:, 8 * kmer_words:] = counts.view(np.uint8).reshape(n, -1)
    return rows


def main():
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libkmdiff_ref.so missing: run `make -C oracle ref` first")
    rng = np.random.default_rng(20260117)

    # ---- the reference's own unit-test goldens -------------------------------------------------
    ref_tests = {
        "source": "tlemane/kmdiff tests/{factorial,model,corrector}_test.cpp + SURVEY.md §8c known answers",
        "lf_table": {"size": 50, "values": {"0": 0.0, "1": 0.0, "10": 15.104412573075514, "50": 148.47776695177302,
                                            "51": 152.40959258449737, "100": 363.7393755555635}},
        "model_signs": {"nb": 30, "totals": 1, "preload": 10,
                        "cases": [{"controls": 200, "cases": 100, "sign": "CONTROL"},
                                  {"controls": 100, "cases": 200, "sign": "CASE"},
                                  {"controls": 100, "cases": 100, "sign": "NO"}]},
        "correctors": [
            {"type": "NOTHING", "threshold": 0.05, "total": 0, "prime": 0, "accept": 0.04, "reject": 0.06},
            {"type": "BONFERRONI", "threshold": 0.05, "total": 100, "prime": 0, "accept": 0.0004, "reject": 0.0006},
            {"type": "BENJAMINI", "threshold": 0.25, "total": 25, "prime": 0, "accept": 0.009, "reject": 0.02},
            {"type": "SIDAK", "threshold": 0.05, "total": 100, "prime": 0, "accept": 0.00050, "reject": 0.00052},
            {"type": "HOLM", "threshold": 0.05, "total": 100, "prime": 90, "accept": 0.004, "reject": 0.006}],
        "known_answers": [
            {"nc": 30, "nk": 30, "totals_c": [1] * 30, "totals_k": [1] * 30, "preload": 10000,
             "controls": [200] * 30, "cases": [100] * 30,
             "p": 1.0932047323640264e-223, "sign": 0, "mean_control": 6000.0, "mean_case": 3000.0},
            {"nc": 10, "nk": 10, "totals_c": [10**9] * 10, "totals_k": [12 * 10**8] * 10, "preload": 10000,
             "controls": [5, 3, 0, 7, 2, 4, 6, 1, 0, 3], "cases": [30, 25, 41, 0, 33, 28, 19, 45, 22, 37],
             "p": 1.7928739662752438e-42, "sign": 1, "mean_control": 37.200000000000003, "mean_case": 280.0}],
    }
    # re-check the transcription against the real thing before writing it
    lf = oracle.ref_lf(50, [0, 1, 10, 50, 51, 100])
    assert list(lf) == [ref_tests["lf_table"]["values"][k] for k in ("0", "1", "10", "50", "51", "100")]
    for ka in ref_tests["known_answers"]:
        row = np.zeros(8 + 4 * (ka["nc"] + ka["nk"]), np.uint8)
        row[8:] = np.array(ka["controls"] + ka["cases"], np.uint32).view(np.uint8)
        p, s, mc, kk = oracle.ref_model_rows(ka["nc"], ka["nk"], ka["totals_c"] + ka["totals_k"], ka["preload"], row, 1, 4)
        assert (p[0], s[0], mc[0], kk[0]) == (ka["p"], ka["sign"], ka["mean_control"], ka["mean_case"]), (p, s, mc, kk)
    for c in ref_tests["correctors"]:
        rc = oracle.RefCorrector(oracle.CORRECTIONS[c["type"].lower()], c["threshold"], c["total"])
        for _ in range(c["prime"]):
            rc.apply(0.0)
        assert rc.apply(c["accept"]) and not rc.apply(c["reject"]), c
    with open(os.path.join(HERE, "reference_tests.json"), "w") as f:
        json.dump(ref_tests, f, indent=1)

    # ---- vectors from the reference library -------------------------------------------------------
    vec = {"source": "oracle/_ref/libkmdiff_ref.so (unmodified tlemane/kmdiff @ /root/reference), x86-64 glibc"}
    idx = sorted(set([0, 1, 2, 3, 9999, 10000, 10001, 12345, 65535, 200000] + rng.integers(0, 30000, 40).tolist()))
    vec["lf"] = {"size": 10000, "idx": idx, "values": hx(oracle.ref_lf(10000, idx))}

    xs = np.concatenate([[0.0, 1e-300, 1e-12, 0.5, 0.999999, 1.0, 1.0000001, 2.0, 25.2, 50.0, 700.0, 1400.0, 1419.5, 1500.0, 1e6],
                         np.exp(rng.uniform(np.log(1e-6), np.log(2000.0), 120))])
    vec["chisqc"] = {"v": 1.0, "x": hx(xs), "p": hx([oracle.ref_chisqc(1.0, float(x)) for x in xs])}

    cohorts = []
    for name, nc, nk, k, b, n, lfsz in [("c2", 10, 10, 31, 4, 160, 10000), ("b2", 6, 9, 63, 2, 120, 10000),
                                        ("b1", 12, 20, 21, 1, 120, 500), ("tiny", 1, 1, 20, 4, 80, 10000),
                                        ("wide", 40, 40, 31, 4, 60, 10000)]:
        kw = (k + 31) // 32
        rows = cohort_rows(rng, nc, nk, b, n, kw)
        totals = rng.integers(5 * 10**8, 15 * 10**8, size=nc + nk).astype(np.uint64)
        p, s, mc, kk = oracle.ref_model_rows(nc, nk, totals, lfsz, rows, kw, b)
        cohorts.append({"name": name, "nc": nc, "nk": nk, "k": k, "count_bytes": b, "lf_size": lfsz,
                        "totals": [int(t) for t in totals], "rows_hex": rows.tobytes().hex(), "n_rows": n,
                        "p": hx(p), "sign": [int(x) for x in s], "mean_control": hx(mc), "mean_case": hx(kk)})
    vec["cohorts"] = cohorts

    corr = []
    ps = np.sort(np.exp(rng.uniform(np.log(1e-14), np.log(1e-3), 300)))
    for kind in ("nothing", "bonferroni", "benjamini", "sidak", "holm"):
        for total in (300, 10**6, 10**9):
            thr = 1e-6 if kind == "nothing" else 0.05
            rc = oracle.RefCorrector(oracle.CORRECTIONS[kind], thr, total)
            corr.append({"type": kind, "threshold": thr, "total": total, "p": hx(ps),
                         "apply_in_order": [int(rc.apply(float(x))) for x in ps]})
    vec["correctors"] = corr
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(vec, f)
    print("wrote", os.path.join(HERE, "reference_tests.json"), os.path.join(HERE, "ref_vectors.json"))


if __name__ == "__main__":
    main()
