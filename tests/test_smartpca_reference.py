"""Pins the oracle's restatement of HAWK's smartpca (the external program `kmdiff diff --pop-correction` runs) and
of the whole population-stratification pipeline:

  * tests/golden/smartpca_vectors.json — pcs.evec / eigenvalues written by the UNMODIFIED smartpca + evec2pca.perl
    (built from the reference's vendored sources, generator tests/golden/make_smartpca_golden.py);
  * where oracle/_ref/{kmdiff,smartpca} exist — the reference binary's own `--pop-correction` run: its .geno file
    against the oracle's Sampler, its pcs.evec against the oracle's PCA, the GLM p-values of its spill files
    against oracle.PopStrat, bit for bit.

Eigenvector signs are LAPACK's choice (dspev) and carry no meaning: PC columns are compared up to sign.
The product's PCA (kmd_smartpca_pcs) is held to the same vectors."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from kmdiff_b200 import build as kb
from kmdiff_b200 import kmtricks_io as kio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    with open(os.path.join(GOLDEN, "smartpca_vectors.json")) as f:
        return json.load(f)["cases"]


def columns_equal_up_to_sign(got, want):
    """-> number of columns that are the exact negation (the rest must be exactly equal)"""
    flipped = 0
    for j in range(want.shape[1]):
        if np.array_equal(got[:, j], want[:, j]):
            continue
        assert np.array_equal(got[:, j], -want[:, j]), f"PC {j + 1} differs beyond its sign"
        flipped += 1
    return flipped


@pytest.mark.parametrize("ci", range(4))
def test_oracle_and_product_pca_match_real_smartpca(ci):
    c = load_cases()[ci]
    m, n = c["m"], c["n"]
    geno = np.unpackbits(np.frombuffer(bytes.fromhex(c["geno_bits"]), dtype=np.uint8))[:m * n].reshape(m, n)
    want = np.array([[float(x) for x in line.split()] for line in c["pcs_evec"].split("\n") if line.strip()])
    assert want.shape == (n, 10)
    xtx, used = oracle.smartpca_xtx(geno, c["diploid"])
    assert used == int((geno.sum(axis=1) > 0).sum())
    evals, evecs = oracle.sym_eig(xtx)
    assert np.allclose(evals[:len(c["evals"])], c["evals"], rtol=0, atol=6e-7)      # smartpca prints "%12.6f"
    pcs = oracle.smartpca_pcs(geno, evecs, 10, c["diploid"])
    columns_equal_up_to_sign(pcs, want)
    # the product's host-side PCA (kmd_smartpca_pcs) on the same matrix
    ev2, pcs2 = K.smartpca_pcs(xtx, 10)
    assert np.allclose(ev2[:len(c["evals"])], c["evals"], rtol=0, atol=6e-7)
    columns_equal_up_to_sign(pcs2, want)


def _structured_cohort(nc, nk, m, seed):
    rng = np.random.default_rng(seed)
    n = nc + nk
    pop = (rng.random(n) < 0.4).astype(np.float64)
    lam = rng.gamma(2, 3, size=(m, 1)) * (1 + 0.8 * pop[None, :] * (rng.random((m, 1)) < 0.3))
    counts = rng.poisson(lam).astype(np.uint32)
    counts[rng.random(m) < 0.02, nc:] += rng.poisson(40, size=nk).astype(np.uint32)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    return counts


@pytest.mark.skipif(not (os.path.exists(oracle.REF_BIN) and os.path.exists(oracle.REF_SMARTPCA)),
                    reason="oracle/_ref/kmdiff + smartpca not built (no /root/reference)")
@pytest.mark.parametrize("nc,nk,npc,seed", [(30, 30, 4, 31), (16, 24, 10, 0)])
def test_reference_pop_correction_pipeline_equals_oracle_pipeline(tmp_path, nc, nk, npc, seed):
    """the reference binary end to end (`-t 1`, one partition so the visiting order is fixed) vs the oracle"""
    n, m, kmer_pca = nc + nk, 40_000, 0.05
    counts = _structured_cohort(nc, nk, m, seed=5 + nc)
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(977) + np.uint64(5)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    spec = K.SynthSpec(nc, nk, kmer_size=31, count_bytes=4, total_rows=m, seed=1)
    d, out = str(tmp_path / "run"), str(tmp_path / "out")
    kio.write_run_dir(d, spec, m, 1, payload=rows.ravel(), totals=totals)
    r = subprocess.run([oracle.REF_BIN, "diff", "--km-run", d, "-1", str(nc), "-2", str(nk), "-o", out, "--pop-correction",
                        "--kmer-pca", str(kmer_pca), "--random-seed", str(seed), "--n-pc", str(npc), "--keep-tmp", "-s", "0.5",
                        "-u", "100000", "-t", "1"], capture_output=True, text=True, timeout=900)
    assert "[error]" not in r.stderr and "Population correction done." in r.stderr, r.stderr[-800:]
    pop = os.path.join(out, "popstrat")

    # 1. Sampler: the .geno file is the presence vectors of the rows the oracle's Sampler picks, in row order
    picked = oracle.sampler_rows(seed, kmer_pca, m).astype(bool)
    want_geno = (counts[picked] > 0).astype(np.uint8)
    with open(os.path.join(pop, "gwas_eigenstratX.geno")) as f:
        ref_geno = np.array([[1 if t == "1" else 0 for t in line.rstrip("\n").split("\t") if t != ""] for line in f], dtype=np.uint8)
    assert ref_geno.shape == want_geno.shape and np.array_equal(ref_geno, want_geno)

    # 2. PCA: pcs.evec against the oracle's smartpca restatement
    Zref = np.loadtxt(os.path.join(pop, "pcs.evec"))
    xtx, _ = oracle.smartpca_xtx(want_geno, True)
    _, evecs = oracle.sym_eig(xtx)
    columns_equal_up_to_sign(oracle.smartpca_pcs(want_geno, evecs, 10, True), Zref)

    # 3. LRT prefilter + GLM: the spill files before / after the stratification correction
    want, _, _ = oracle.Model(nc, nk, totals).diff(rows.ravel(), 1, 4, 0.5 / 100000)
    rec = 8 + 28 + 2 + 8 * n
    cli = kb.build_cli()
    got = {}
    for suffix in ("_uncorrected", "_popstrat_uncorrected"):
        raw = subprocess.run([cli, "dev-cat", os.path.join(out, "partitions", f"p0{suffix}"), "0", "lz4"], capture_output=True, timeout=60).stdout
        assert len(raw) % rec == 0 and len(raw) // rec == len(want)
        a = np.frombuffer(raw, dtype=np.uint8).reshape(-1, rec)
        got[suffix] = (np.ascontiguousarray(a[:, :8]).view("<u8").ravel(), np.ascontiguousarray(a[:, 8:16]).view("<f8").ravel(),
                       np.ascontiguousarray(a[:, 38:]).view("<f8").reshape(-1, n))
    km, p_lrt, cnt = got["_uncorrected"]
    assert np.array_equal(km, want["kmer"][:, 0]) and np.array_equal(p_lrt, want["p"])
    assert np.array_equal(cnt, counts[want["row"]].astype(np.float64))
    ps = oracle.PopStrat(nc, nk, totals, Zref, npc)
    p_glm = np.array([ps.apply(c) for c in cnt])
    assert np.array_equal(got["_popstrat_uncorrected"][1], p_glm)          # bit for bit
    assert len(want) > 100
