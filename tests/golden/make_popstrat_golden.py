#!/usr/bin/env python
"""Generates tests/golden/popstrat_vectors.json from the UNMODIFIED reference
(oracle/_ref/libkmdiff_ref.so: libstdc++'s engine behind kmdiff::Sampler, src/linear_model.cpp,
src/popstrat.cpp).  Run in the build container only:

    python tests/golden/make_popstrat_golden.py

Doubles are stored as C99 hex-float strings (bit-exact).  smartpca is NOT covered: the reference's
smartpca cannot be built in this image (GSL/LAPACKE/OpenBLAS headers absent), see
oracle/kmd_popstrat_oracle.h.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def ref_sampler(seed, v, n):
    out = np.zeros(n, np.uint8)
    oracle.ref().ref_sampler_rows(seed, v, n, out.ctypes.data)
    return np.flatnonzero(out).tolist()


def cohort_case(rng, nc, nk, npc, n_kmers):
    n = nc + nk
    totals = rng.integers(int(8e8), int(1.2e9), n).astype(np.uint64)
    # stratified cohort: two sub-populations, PCs quantised like pcs.evec (%.4f)
    pop = rng.random(n) < 0.4
    Z = rng.normal(scale=0.05, size=(n, 10))
    Z[:, 0] += np.where(pop, 0.08, -0.05)
    Z = np.round(Z, 4)
    counts = []
    for t in range(n_kmers):
        lam = rng.gamma(2.0, 3.0)
        c = rng.poisson(lam, n).astype(np.float64)
        kind = t % 4
        if kind == 1:
            c[nc:] += rng.poisson(25.0, nk)           # case-enriched
        elif kind == 2:
            c[pop] += rng.poisson(25.0, int(pop.sum()))  # population-specific, not phenotype
        elif kind == 3:
            c[:nc] += rng.poisson(15.0, nc)           # control-enriched
        counts.append(c)
    return totals, Z, np.array(counts)


def main():
    rng = np.random.default_rng(20261017)
    out = {"sampler": [], "irls": [], "inverse": [], "popstrat": []}
    for seed, v, n in ((0, 0.01, 100000), (42, 0.001, 300000), (2147483647, 0.5, 200), (123456789012, 0.05, 5000)):
        out["sampler"].append({"seed": seed, "kmer_pca": v, "n": n, "taken": ref_sampler(seed, v, n)})

    for n, p, kind in ((40, 4, "plain"), (60, 13, "plain"), (30, 3, "separable"), (20, 3, "collinear"), (50, 6, "scaled")):
        x = np.c_[np.ones(n), rng.normal(size=(n, p - 1))]
        beta = rng.normal(size=p)
        y = (rng.random(n) < 1 / (1 + np.exp(-(x @ beta)))).astype(np.float64)
        if kind == "separable":
            y = (x[:, 1] > 0).astype(np.float64)
        if kind == "collinear":
            x[:, 2] = 2.0 * x[:, 1]
        if kind == "scaled":
            x[:, 1] *= 1e9
            x[:, 2] *= 1e-8
        w, s, q, e, it = oracle.glm_irls(x, y, impl="ref")
        out["irls"].append({"kind": kind, "n": n, "p": p, "x": hx(x), "y": hx(y), "weights": hx(w),
                            "singular": s, "nan": q, "error": float(e).hex(), "iters": it})

    for n in (2, 5, 13):
        a = rng.normal(size=(n, n))
        a = a @ a.T + n * np.eye(n)
        inv, s, q = oracle.inverse(a, impl="ref")
        out["inverse"].append({"n": n, "a": hx(a), "inv": hx(inv), "singular": s, "nan": q})

    with tempfile.TemporaryDirectory() as d:
        for nc, nk, npc, nk_mers in ((12, 12, 10, 24), (20, 30, 2, 16), (40, 40, 10, 12)):
            totals, Z, counts = cohort_case(rng, nc, nk, npc, nk_mers)
            rp = oracle.RefPopStrat(nc, nk, totals, Z, npc, d)
            p = [rp.apply(c) for c in counts]
            out["popstrat"].append({"nb_controls": nc, "nb_cases": nk, "npc": npc,
                                    "totals": [int(t) for t in totals], "Z": hx(Z),
                                    "null_features": hx(rp.null_features), "alt_features": hx(rp.alt_features),
                                    "null_model": hx(rp.null_model), "counts": hx(counts), "p": hx(p)})
    with open(os.path.join(HERE, "popstrat_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote popstrat_vectors.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
