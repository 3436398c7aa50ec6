"""GPU parity tests of the --pop-correction path through the C-ABI, against
oracle/kmd_popstrat_oracle.c on the same seeded inputs.

  * Sampler + presence emit: the sampled row set and the 0/1 vectors are bit-exact.
  * Gram matrix (int8 tcgen05 GEMM per carrier-count bucket + fp64 fix-up): the integer part is exact;
    the fp64 combination groups the reference's per-SNP sum by bucket, so XTX agrees to ~1e-13 relative
    (tolerance 1e-10), and the "%10.4f" principal components within one quantum.
  * GLM: iteration counts identical, weights / LLR / p-values within GLM_RTOL of the oracle (the fit
    follows the reference's summation order; the only deviation is CUDA's pow/log vs glibc's).
"""
import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from conftest import SHAPES, make_cohort, rel_err

pytestmark = pytest.mark.gpu

THR = 0.05 / 100000
GLM_RTOL = 1e-9


def ctx_for(spec, totals, thr=THR, **kw):
    nc, nk = spec.nb_controls, spec.nb_cases
    return K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=spec.kmer_size,
                         count_bytes=spec.count_bytes, threshold=thr, **kw)


def counts_of(spec, rows):
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[spec.count_bytes]
    r = rows.reshape(-1, spec.row_bytes)[:, 8 * spec.kmer_words:]
    return np.ascontiguousarray(r).view(dt).reshape(len(r), -1)


@pytest.mark.parametrize("shape,seed,v", [("c2", 0, 0.01), ("c5", 7, 0.05), ("b1", 2**31 - 1, 0.002), ("odd", 42, 0.3),
                                          ("c4", 12345, 0.02)])
def test_sampler_and_presence(shape, seed, v):
    n = 120_011 if shape != "c4" else 30_003
    spec, rows, totals = make_cohort(shape, n, seed=3)
    base = 2**31 - 50_000      # the draw index crosses the engine's period inside the first batch
    cut = (n // 3) // 16 * 16
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(v, seed=seed, n_pc=2)
        ctx.scan_host(rows[:cut * spec.row_bytes], row_base=base)
        ctx.scan_host(rows[cut * spec.row_bytes:], row_base=base + cut)
        geno, grow = ctx.popstrat_geno()
        surv = ctx.survivors()
        sc = ctx.survivor_counts()
    take = K.sampler_rows(seed, v, base, n).astype(bool)      # pinned against the oracle in test_popstrat_host.py
    idx = np.flatnonzero(take)
    assert len(idx) > 0
    assert np.array_equal(grow, idx.astype(np.uint64) + np.uint64(base))
    cnt = counts_of(spec, rows)
    assert np.array_equal(geno, (cnt[idx] > 0).astype(np.uint8))
    # the survivors keep their counts (KmerSign::m_counts_ratio)
    assert len(surv) > 0
    assert np.array_equal(sc, cnt[(surv["row"] - np.uint64(base)).astype(np.int64)])


def test_sampler_small_index_matches_oracle_on_device():
    spec, rows, totals = make_cohort("c2", 50_000, seed=4)
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(0.07, seed=99, n_pc=2)
        ctx.scan_host(rows)
        _, grow = ctx.popstrat_geno()
    assert np.array_equal(grow, np.flatnonzero(oracle.sampler_rows(99, 0.07, 50_000)).astype(np.uint64))


def test_sampler_overflow_reported():
    spec, rows, totals = make_cohort("c2", 20_000, seed=4)
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(0.5, seed=1, n_pc=2, max_sampled=100)
        ctx.scan_host(rows)
        with pytest.raises(K.KmdError) as e:
            ctx.popstrat_sampled()
        assert e.value.code == 4


WIDE = (300, 330, 31, 1)   # 630 samples: two 512-column halves of the accumulator


@pytest.mark.parametrize("shape,n_rows,v,diploid", [("c2", 200_000, 0.05, True), ("c3", 40_000, 0.25, True),
                                                      ("c4", 30_000, 0.5, True), ("c4", 8_000, 1.0, False),
                                                      ("b1", 100_000, 0.1, False), ("wide", 12_000, 0.5, True),
                                                      ("tiny", 3_000, 1.0, True)])
def test_gram_matches_oracle(shape, n_rows, v, diploid):
    if shape == "wide":
        SHAPES["wide"] = WIDE
    spec, rows, totals = make_cohort(shape, n_rows, seed=8, enriched_ppm=20000)
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(v, seed=5, is_diploid=diploid, n_pc=2)
        ctx.scan_host(rows)
        geno, _ = ctx.popstrat_geno()
        xtx_raw, used_raw = ctx.popstrat_gram(normalize=False)
        xtx, used = ctx.popstrat_gram(normalize=True)
        ms = ctx.popstrat_gram_ms()
    n = spec.nb_controls + spec.nb_cases
    assert geno.shape[0] > 100 or shape == "tiny"
    want, w_used = oracle.smartpca_xtx(geno, is_diploid=diploid)
    assert used == w_used == used_raw
    assert np.array_equal(xtx, xtx.T)
    scale = np.abs(want).max()
    assert np.abs(xtx - want).max() <= 1e-10 * scale
    tr = np.trace(xtx_raw) / (n - 1)
    assert np.allclose(xtx_raw / tr, xtx, rtol=1e-14, atol=0)
    assert ms > 0
    # integer check of the tensor-core part alone: with every weight 1 and mu 0 the formula is G itself;
    # recover G's diagonal-sum identity trace(G) = total carriers from the un-normalised matrix instead
    c = geno.sum(1)
    keep = c > 0
    mu = c[keep] / n
    p = 1 - np.sqrt(1 - mu) if diploid else mu
    y = p * (1 - p)
    f2 = np.where(y > 0, 1 / np.where(y > 0, y, 1), 1.0)
    X = (geno[keep] - mu[:, None])
    ref_raw = (X * f2[:, None]).T @ X
    assert np.abs(xtx_raw - ref_raw).max() <= 1e-10 * np.abs(ref_raw).max()


@pytest.mark.parametrize("shape,n_rows,v,splits", [("c2", 200_000, 0.05, 3), ("c4", 30_000, 0.3, 2), ("odd", 50_000, 0.2, 5)])
def test_sharded_gram_equals_single_context_bit_for_bit(shape, n_rows, v, splits):
    """the multi-GPU Gram exchange without NCCL: the rows are split over several contexts (the sampler draws by global
    row index, so the sampled set is the same), each computes its integer co-carrier matrices per carrier count
    (kmd_popstrat_gram_partial), the integers are summed — what the all-reduce does — and combined in fp64
    (kmd_popstrat_gram_combine): the matrix equals the single-context kmd_popstrat_gram bit for bit, normalised or not"""
    import torch
    spec, rows, totals = make_cohort(shape, n_rows, seed=19)
    R, n = spec.row_bytes, spec.nb_controls + spec.nb_cases
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(v, seed=5, n_pc=2)
        ctx.scan_host(rows)
        want, used = ctx.popstrat_gram(normalize=False)
        want_n, _ = ctx.popstrat_gram(normalize=True)
        # one context through the partial / combine pair: same bits
        G = torch.empty(n * n * n, dtype=torch.int32, device="cuda")
        H = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.popstrat_gram_partial(G.data_ptr(), H.data_ptr())
        got, used1 = ctx.popstrat_gram_combine(G.data_ptr(), H.data_ptr(), normalize=False)
        assert used1 == used > 100 and np.array_equal(got, want)
        assert int(H.sum()) >= used and int(G.view(n, n, n).diagonal(dim1=1, dim2=2).sum()) > 0
    cuts = [0] + sorted(int(x) // 16 * 16 for x in np.random.default_rng(1).integers(1, n_rows, splits - 1)) + [n_rows]
    Gs = torch.zeros(n * n * n, dtype=torch.int32, device="cuda")
    Hs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    last = None
    for a, b in zip(cuts[:-1], cuts[1:]):
        c2 = ctx_for(spec, totals)
        c2.popstrat_enable(v, seed=5, n_pc=2)
        if b > a:
            c2.scan_host(rows[a * R:b * R], row_base=a)
        Gp = torch.empty_like(Gs)
        Hp = torch.empty_like(Hs)
        torch.cuda.synchronize()
        c2.popstrat_gram_partial(Gp.data_ptr(), Hp.data_ptr())
        Gs += Gp
        Hs += Hp
        if last is not None:
            last.close()
        last = c2
    torch.cuda.synchronize()
    got, used2 = last.popstrat_gram_combine(Gs.data_ptr(), Hs.data_ptr(), normalize=False)
    got_n, _ = last.popstrat_gram_combine(Gs.data_ptr(), Hs.data_ptr(), normalize=True)
    last.close()
    assert used2 == used
    assert np.array_equal(got, want) and np.array_equal(got_n, want_n)


def test_gram_empty_and_all_present():
    spec, rows, totals = make_cohort("c2", 5_000, seed=8)
    with ctx_for(spec, totals) as ctx:
        ctx.popstrat_enable(0.0, seed=5, n_pc=2)
        ctx.scan_host(rows)
        assert ctx.popstrat_sampled() == 0
        xtx, used = ctx.popstrat_gram()
    assert used == 0 and not xtx.any()


def _cohort_with_structure(nc, nk, m, seed):
    """counts with a sub-population signal so the leading PCs are well separated"""
    rng = np.random.default_rng(seed)
    n = nc + nk
    pop = (rng.random(n) < 0.4).astype(np.float64)
    lam = rng.gamma(2, 3, size=(m, 1)) * (1 + 0.8 * pop[None, :] * (rng.random((m, 1)) < 0.3))
    counts = rng.poisson(lam).astype(np.uint32)
    counts[rng.random(m) < 0.02, nc:] += rng.poisson(40, size=nk).astype(np.uint32)
    totals = rng.integers(int(5e8), int(2e9), n).astype(np.uint64)
    return counts, totals


@pytest.mark.parametrize("nc,nk,npc", [(25, 35, 10), (100, 100, 2), (200, 200, 10), (12, 12, 0)])
def test_glm_matches_oracle(nc, nk, npc):
    n = nc + nk
    rng = np.random.default_rng(nc + npc)
    counts, totals = _cohort_with_structure(nc, nk, 150 if n <= 200 else 60, seed=nc)
    counts[0] = 0                    # a k-mer nobody carries
    counts[1] = counts[1].max()      # constant feature column
    Z = np.round(rng.normal(scale=0.1, size=(n, 10)), 4)
    ps = oracle.PopStrat(nc, nk, totals, Z, npc)
    with K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), count_bytes=4) as ctx:
        ctx.popstrat_enable(0.0, n_pc=npc)
        null_model = ctx.popstrat_init(Z, totals)
        p, model, llr, iters = ctx.popstrat_fit(counts)
    assert rel_err(null_model, ps.null_model).max() <= GLM_RTOL
    worst = 0.0
    for i, c in enumerate(counts):
        wp, wmodel, wit, wllr = ps.apply(c.astype(np.float64), full=True)
        assert iters[i] == wit, (i, iters[i], wit)
        if np.isnan(wp):
            assert np.isnan(p[i])
            continue
        assert rel_err(model[i], wmodel).max() <= 1e-6, (i, model[i], wmodel)
        assert abs(llr[i] - wllr) <= GLM_RTOL * max(abs(wllr), 1.0), (i, llr[i], wllr)
        if wp > 0:
            # d ln p = -dLLR/2 for large LLR: the stated 1e-9 bound on the statistic carries over
            tol = GLM_RTOL * max(1.0, abs(wllr))
            assert abs(p[i] - wp) <= tol * wp, (i, p[i], wp)
            worst = max(worst, abs(p[i] - wp) / wp)
        else:
            assert p[i] == 0
    print(f"glm {nc}+{nk} npc={npc}: max rel p err {worst:.2e}")


def test_row_order_survives_a_p_sort_followed_by_the_glm():
    """sorted_pvalues() / a BENJAMINI correct() leave the survivor slots in p order; popstrat_apply then rewrites the
    p-values, which invalidates that order — survivors() and survivor_counts() must come back in ROW order again, and a
    later p-sort must break ties by row, not by the stale order"""
    nc, nk, npc, m = 30, 30, 3, 40_000
    n = nc + nk
    counts, _ = _cohort_with_structure(nc, nk, m, seed=5)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(13) + np.uint64(1)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    with K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), count_bytes=4, threshold=THR) as ctx:
        ctx.popstrat_enable(0.2, seed=3, n_pc=npc)
        ctx.scan_host(rows.ravel())
        first = ctx.survivors()
        assert len(first) > 20 and np.all(np.diff(first["row"].astype(np.int64)) > 0)
        xtx, _ = ctx.popstrat_gram()
        _, Z = K.smartpca_pcs(xtx, 10)
        ctx.popstrat_init(Z, totals)
        ps = ctx.sorted_pvalues()                                   # slots now in p order
        assert np.all(np.diff(ps) >= 0)
        ctx.correct(K.CorrectionType.BENJAMINI, 0.05, m)
        ctx.popstrat_apply()                                        # p-values replaced
        after = ctx.survivors()
        sc = ctx.survivor_counts()
        assert np.array_equal(after["row"], first["row"])          # row order, not the stale p order
        assert np.array_equal(sc, counts[first["row"].astype(np.int64)])
        assert not np.array_equal(after["p_value"], first["p_value"])
        ps2 = ctx.sorted_pvalues()
        assert np.array_equal(ps2, np.sort(after["p_value"]))


def test_pop_correction_end_to_end():
    """scan (+sampler) -> Gram -> PCs -> null model -> GLM on every survivor -> Bonferroni, against the
    same pipeline assembled from oracle pieces."""
    nc, nk, npc, m = 40, 40, 4, 60_000
    n = nc + nk
    counts, _ = _cohort_with_structure(nc, nk, m, seed=77)
    counts[:, 0] = np.maximum(counts[:, 0], 1)     # kmtricks never emits an all-zero row
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(977) + np.uint64(5)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    v, seed = 0.2, 31
    with K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), count_bytes=4, threshold=THR) as ctx:
        ctx.popstrat_enable(v, seed=seed, n_pc=npc)
        ctx.scan_host(rows.ravel())
        before = ctx.survivors()
        geno, _ = ctx.popstrat_geno()
        xtx, used = ctx.popstrat_gram()
        evals, Z = K.smartpca_pcs(xtx, 10)
        ctx.popstrat_init(Z, totals)
        ctx.popstrat_apply()
        after = ctx.survivors()
        n_ctl, n_case = ctx.correct(K.CorrectionType.BONFERRONI, 0.05, m)
        sig = ctx.significant()
    # oracle side
    take = oracle.sampler_rows(seed, v, m)
    assert np.array_equal(geno, (counts[take] > 0).astype(np.uint8))
    oxtx, oused = oracle.smartpca_xtx(geno)
    assert used == oused
    oe, ov = oracle.sym_eig(oxtx)
    oZ = oracle.smartpca_pcs(geno, ov, 10)
    assert np.abs(Z - oZ).max() <= 1e-4 + 1e-12       # one "%10.4f" quantum at most
    assert np.mean(Z != oZ) < 0.02
    ps = oracle.PopStrat(nc, nk, totals, Z, npc)     # same Z on both sides from here on
    model = oracle.Model(nc, nk, totals)
    want, _, _ = model.diff(rows.ravel(), 1, 4, THR)
    assert np.array_equal(before["row"], want["row"]) and len(want) > 20
    assert np.array_equal(after["row"], before["row"]) and np.array_equal(after["sign"], before["sign"])
    wp = np.array([ps.apply(counts[r].astype(np.float64)) for r in want["row"].astype(np.int64)])
    ok = wp > 0
    assert np.array_equal(after["p_value"][~ok], wp[~ok])
    assert (np.abs(after["p_value"][ok] - wp[ok]) <= 1e-7 * wp[ok]).all()
    keep = oracle.aggregate(wp, oracle.BONFERRONI, 0.05, m)
    edge = np.abs(wp - 0.05 / m) <= 1e-7 * (0.05 / m)
    if not edge.any():
        assert np.array_equal(sig["row"], want["row"][keep])
        assert n_ctl + n_case == len(keep)


@pytest.mark.parametrize("lz4", [False, True])
def test_cli_pop_correction(tmp_path, lz4):
    """`kmdiff-b200 diff --pop-correction` on a run directory == the same pipeline driven through the
    C-ABI in process (which the tests above hold against the oracle): popstrat/ files, corrected
    spill files and the FASTA output."""
    import os
    import subprocess

    from kmdiff_b200 import build as kb
    from kmdiff_b200 import kmtricks_io as kio
    from test_host_cli import read_fasta, strip_index

    cli = kb.build_cli()
    nc, nk, npc, m, parts = 40, 40, 4, 60_000, 4
    n = nc + nk
    counts, _ = _cohort_with_structure(nc, nk, m, seed=78)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(977) + np.uint64(5)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    # -s 0.5: with 80 samples the GLM p-values saturate near 2e-6 (perfect separation), above 0.05/m
    v, seed = 0.2, 31
    spec = K.SynthSpec(nc, nk, kmer_size=31, count_bytes=4, total_rows=m, seed=1)
    d, out = str(tmp_path / "run"), str(tmp_path / "out")
    kio.write_run_dir(d, spec, m, parts, lz4=lz4, payload=rows.ravel(), totals=totals)
    r = subprocess.run([cli, "diff", "-d", d, "-1", str(nc), "-2", str(nk), "-o", out, "--pop-correction", "--kmer-pca", str(v),
                        "--random-seed", str(seed), "--n-pc", str(npc), "--keep-tmp", "-c", "bonferroni", "-s", "0.5",
                        "-u", "100000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "[error]" not in r.stderr, r.stderr
    assert "PCA done." in r.stderr and "Population correction done." in r.stderr

    with K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), count_bytes=4, threshold=0.5 / 100000) as ctx:
        ctx.popstrat_enable(v, seed=seed, n_pc=npc)
        ctx.scan_host(rows.ravel())
        before = ctx.survivors()
        geno, grow = ctx.popstrat_geno()
        xtx, used = ctx.popstrat_gram()
        evals, Z = K.smartpca_pcs(xtx, 10)
        ctx.popstrat_init(Z, totals)
        ctx.popstrat_apply()
        after = ctx.survivors()
        ctx.correct(K.CorrectionType.BONFERRONI, 0.5, m)
        sig = ctx.significant()
    assert len(before) > 20 and len(sig) > 0

    pop = os.path.join(out, "popstrat")
    with open(os.path.join(pop, "gwas_eigenstratX.geno")) as f:
        lines = f.read().split("\n")[:-1]
    assert len(lines) == len(geno)
    want_lines = ["".join("1\t" if x else "0\t" for x in g) for g in geno]
    assert lines == want_lines
    with open(os.path.join(pop, "gwas_eigenstratX.snp")) as f:
        snp = f.read().split("\n")[:-1]
    assert snp[0] == "0\t1\t0.0\t0" and snp[-1] == f"{len(geno) - 1}\t1\t0.0\t0" and len(snp) == len(geno)
    Zf = np.loadtxt(os.path.join(pop, "pcs.evec"))
    assert Zf.shape == (n, 10) and np.abs(Zf - Z).max() < 1e-12
    with open(os.path.join(pop, "gwas_eigenstratX.ind")) as f:
        ind = f.read().split("\n")[:-1]
    assert ind[0] == "C0\tU\tControl" and ind[-1] == f"K{nk - 1}\tU\tCase" and len(ind) == n
    with open(os.path.join(pop, "gwas_eigenstratX.total")) as f:
        assert [int(x) for x in f.read().split()] == [int(t) for t in totals]

    # spill files: records | u16 N | N x f64 counts; p{i}_uncorrected holds the LRT p, p{i}_popstrat_uncorrected the GLM p
    rec = 8 + 28 + 2 + 8 * n
    for suffix, want in (("_uncorrected", before), ("_popstrat_uncorrected", after)):
        got_p, got_counts = [], []
        for p in range(parts):
            raw = subprocess.run([cli, "dev-cat", os.path.join(out, "partitions", f"p{p}{suffix}"), "0", "lz4"],
                                 capture_output=True, timeout=60).stdout
            assert len(raw) % rec == 0
            for i in range(len(raw) // rec):
                b = raw[i * rec:(i + 1) * rec]
                got_p.append(np.frombuffer(b[8:16], dtype=np.float64)[0])
                assert np.frombuffer(b[36:38], dtype=np.uint16)[0] == n
                got_counts.append(np.frombuffer(b[38:], dtype=np.float64))
        assert len(got_p) == len(want)
        assert np.array_equal(np.array(got_p), want["p_value"], equal_nan=True)
        assert np.array_equal(np.array(got_counts), counts[want["row"].astype(np.int64)].astype(np.float64))

    want_recs = sorted((strip_index(K.format_fasta(0, s, 31).split("\n", 1)[0]),
                        K.format_fasta(0, s, 31).split("\n", 1)[1].replace("\n", "")) for s in sig)
    got = sorted((strip_index(h), s) for f in ("control_kmers.fasta", "case_kmers.fasta")
                 for h, s in read_fasta(os.path.join(out, f)))
    assert got == want_recs


def test_cli_pop_correction_from_count_files(tmp_path):
    """--pop-correction on a run directory of per-sample count files (device merge) leaves the same popstrat/
    files, spill files and FASTA records as on the matrices of the same cohort: the Sampler draws follow the
    merged row numbering."""
    import filecmp
    import os
    import subprocess

    from kmdiff_b200 import build as kb
    from kmdiff_b200 import kmtricks_io as kio

    cli = kb.build_cli()
    nc, nk, npc, m, parts = 24, 24, 3, 40_000, 3
    n = nc + nk
    counts, _ = _cohort_with_structure(nc, nk, m, seed=5)
    counts[:, 0] = np.maximum(counts[:, 0], 1)            # no all-zero row: the merge cannot produce one
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(977) + np.uint64(5)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    spec = K.SynthSpec(nc, nk, kmer_size=31, count_bytes=4, total_rows=m, seed=1)
    outs = {}
    for mode in ("matrix", "counts"):
        d, out = str(tmp_path / f"run_{mode}"), str(tmp_path / f"out_{mode}")
        kio.write_run_dir(d, spec, m, parts, lz4=True, payload=rows.ravel(), totals=totals, count_files=(mode == "counts"))
        r = subprocess.run([cli, "diff", "-d", d, "-1", str(nc), "-2", str(nk), "-o", out, "--pop-correction", "--kmer-pca", "0.2",
                            "--random-seed", "7", "--n-pc", str(npc), "--keep-tmp", "-s", "0.5", "-u", "100000",
                            "--merge-records", "300000"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "[error]" not in r.stderr, r.stderr
        assert "Population correction done." in r.stderr
        outs[mode] = out
    for rel in ("popstrat/gwas_eigenstratX.geno", "popstrat/gwas_eigenstratX.snp", "popstrat/pcs.evec",
                "control_kmers.fasta", "case_kmers.fasta"):
        a, b = os.path.join(outs["matrix"], rel), os.path.join(outs["counts"], rel)
        assert os.path.getsize(a) > 0 and filecmp.cmp(a, b, shallow=False), rel
    for p in range(parts):
        for suffix in ("_uncorrected", "_popstrat_uncorrected"):
            raw = [subprocess.run([cli, "dev-cat", os.path.join(outs[mode], "partitions", f"p{p}{suffix}"), "0", "lz4"],
                                  capture_output=True, timeout=60).stdout for mode in ("matrix", "counts")]
            assert raw[0] == raw[1] and len(raw[0]) > 0


def test_cli_pop_correction_matches_reference_binary():
    """`kmdiff-b200 diff --pop-correction` against the reference binary's own run (oracle/_ref/kmdiff + HAWK smartpca +
    evec2pca.perl) on the same run directory: identical .geno file, PCs equal up to eigenvector sign, identical
    survivor k-mers, LRT and GLM p-values within 1e-9 relative, identical FASTA k-mer sets."""
    import json
    import os
    import subprocess
    import sys

    if not (os.path.exists(oracle.REF_BIN) and os.path.exists(oracle.REF_SMARTPCA)):
        pytest.skip("oracle/_ref/kmdiff + smartpca not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "scripts", "popstrat_vs_reference.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().split("\n")[-1])
    assert res["geno_identical"] and res["pcs_equal_up_to_sign"]
    assert res["kmers_identical_uncorrected"] and res["kmers_identical_popstrat_uncorrected"]
    assert res["max_rel_p_uncorrected"] <= 1e-9 and res["max_rel_p_popstrat_uncorrected"] <= 1e-9
    assert res["fasta_kmers_identical_control_kmers.fasta"] and res["fasta_kmers_identical_case_kmers.fasta"]
