"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same
seeded inputs.  Integer/byte/index results (rows, k-mers, signs, counts, sums) must be bit-exact;
p-values and the two "means" are fp64 and must agree within 1e-9 relative (BASELINE.json
north_star), in practice ~1e-13 (CUDA vs glibc log/exp).
"""
import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from conftest import SHAPES, make_cohort, rel_err
from kmdiff_b200 import _lib as _kmd_lib

K._lib = _kmd_lib

pytestmark = pytest.mark.gpu

P_RTOL = 1e-9   # stated tolerance on fp64 p-values (north_star)
THR = 0.05 / 100000


def ctx_for(spec, totals, thr=THR, **kw):
    nc, nk = spec.nb_controls, spec.nb_cases
    return K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=spec.kmer_size,
                         count_bytes=spec.count_bytes, threshold=thr, **kw)


def oracle_for(spec, totals):
    return oracle.Model(spec.nb_controls, spec.nb_cases, totals)


def check_survivors(got, want):
    assert len(got) == len(want)
    assert np.array_equal(got["row"], want["row"])
    assert np.array_equal(got["kmer"], want["kmer"])
    assert np.array_equal(got["sign"], want["sign"])
    assert np.array_equal(got["mean_case"], want["mean_case"])
    assert np.array_equal(got["mean_control"], want["mean_control"])
    if len(want):
        nz = want["p"] > 0
        assert np.array_equal(got["p_value"][~nz], want["p"][~nz])
        assert rel_err(got["p_value"][nz], want["p"][nz]).max() <= P_RTOL


@pytest.mark.parametrize("shape", list(SHAPES))
def test_scan_matches_oracle(shape):
    # sizes chosen so the ring sees several full tiles per CTA *and* a ragged tail
    n = 300_007 if SHAPES[shape][0] + SHAPES[shape][1] < 100 else 60_013
    spec, rows, totals = make_cohort(shape, n, seed=11)
    m = oracle_for(spec, totals)
    want, w_ctl, w_case = m.diff(rows, spec.kmer_words, spec.count_bytes, THR)
    with ctx_for(spec, totals) as ctx:
        ctx.scan_host(rows)
        c = ctx.counters()
        got = ctx.survivors()
    assert c.rows == n
    assert (c.survivors, c.n_control, c.n_case) == (len(want), w_ctl, w_case)
    assert len(want) > 0
    check_survivors(got, want)


@pytest.mark.parametrize("shape", ["c2", "c5", "b1", "odd", "straddle", "tiny"])
def test_model_rows_every_row(shape):
    """IModel::process on every row (no screen): p, sign, means."""
    spec, rows, totals = make_cohort(shape, 20_000, seed=5)
    p, sign, mc, kk = oracle_for(spec, totals).rows(rows, spec.kmer_words, spec.count_bytes)
    with ctx_for(spec, totals) as ctx:
        gp, gsign, gmc, gkk = ctx.model_rows(rows)
    assert np.array_equal(gsign, sign)
    assert np.array_equal(gkk, kk)
    assert np.array_equal(gmc, mc)
    nz = p > 0
    assert np.array_equal(gp[~nz], p[~nz])
    assert rel_err(gp[nz], p[nz]).max() <= P_RTOL


def test_poisson_likelihood_signs():
    """reference tests/model_test.cpp:45-81 — 30x200 vs 30x100 is CONTROL, swapped CASE, equal NO;
    plus the two known answers measured on the reference."""
    ones = [1] * 30
    m = K.PoissonLikelihood(30, 30, ones, ones, 50)
    p, s, mc, kk = m.process([200] * 30, [100] * 30)
    assert s == K.Significance.CONTROL and (mc, kk) == (6000.0, 3000.0)
    assert abs(p - 1.0932047323640264e-223) / 1.0932047323640264e-223 <= P_RTOL
    assert m.process([100] * 30, [200] * 30)[1] == K.Significance.CASE
    assert m.process([100] * 30, [100] * 30)[1] == K.Significance.NO
    m2 = K.PoissonLikelihood(10, 10, [10**9] * 10, [12 * 10**8] * 10, 10000)
    p, s, mc, kk = m2.process([5, 3, 0, 7, 2, 4, 6, 1, 0, 3], [30, 25, 41, 0, 33, 28, 19, 45, 22, 37])
    assert s == K.Significance.CASE and mc == 37.200000000000003 and kk == 280.0
    assert abs(p - 1.7928739662752438e-42) / 1.7928739662752438e-42 <= P_RTOL


def test_empty_and_tiny_inputs():
    spec, rows, totals = make_cohort("c2", 0)
    with ctx_for(spec, totals) as ctx:
        ctx.scan_host(rows, n_rows=0)
        c = ctx.counters()
        assert (c.rows, c.survivors) == (0, 0)
        assert len(ctx.survivors()) == 0
        assert ctx.correct(K.CorrectionType.BENJAMINI, 0.05, 0) == (0, 0)
        assert len(ctx.significant()) == 0
    # fewer rows than one ring tile: everything goes through the tail path
    for n in (1, 15, 17, 511):
        spec, rows, totals = make_cohort("c2", n, seed=n, enriched_ppm=300_000)
        want, *_ = oracle_for(spec, totals).diff(rows, 1, 4, THR)
        with ctx_for(spec, totals) as ctx:
            ctx.scan_host(rows)
            check_survivors(ctx.survivors(), want)


def test_batches_accumulate_and_reset():
    """partitions arrive as separate scan calls; survivors accumulate across them in row order"""
    spec, rows, totals = make_cohort("c2", 200_000, seed=3)
    want, w_ctl, w_case = oracle_for(spec, totals).diff(rows, 1, 4, THR)
    R = spec.row_bytes
    cuts = [0, 70_001, 70_002, 150_000, 200_000]
    with ctx_for(spec, totals) as ctx:
        for part, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
            ctx.scan_host(rows[lo * R:hi * R], row_base=lo, partition=part)
        got = ctx.survivors()
        check_survivors(got, want)
        part_of = np.searchsorted(cuts, got["row"], side="right") - 1
        assert np.array_equal(got["partition"], part_of)
        ctx.reset()
        assert ctx.counters().rows == 0 and len(ctx.survivors()) == 0
        ctx.scan_host(rows[:1000 * R])
        assert ctx.counters().rows == 1000


def test_device_resident_scan_equals_host_scan():
    import torch
    spec, rows, totals = make_cohort("c5", 123_457, seed=9)
    want, *_ = oracle_for(spec, totals).diff(rows, 2, 2, THR)
    d = torch.from_numpy(rows).cuda()
    with ctx_for(spec, totals) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.scan_device(d.data_ptr(), 123_457)
        check_survivors(ctx.survivors(), want)
        assert ctx.scan_kernel_ms() > 0
        assert ctx.launch_count() >= 2


@pytest.mark.parametrize("shape,buf_rows", [("c2", 4099), ("c3", 1000), ("odd", 77)])
def test_async_host_scan_with_recycled_pinned_buffers(shape, buf_rows):
    """kmd_scan_host_async + kmd_host_wait the way the host program's readers use them: a small pool of pinned buffers
    refilled as soon as their copy is done, buffers submitted out of row order, one waiter thread per buffer"""
    import ctypes
    import threading
    from kmdiff_b200 import _lib
    L = _lib.load()
    spec, rows, totals = make_cohort(shape, 150_001, seed=21)
    R = spec.row_bytes
    want, *_ = oracle_for(spec, totals).diff(rows, spec.kmer_words, spec.count_bytes, THR)
    n_bufs = 3
    ptrs = [L.kmd_host_alloc(buf_rows * R) for _ in range(n_bufs)]
    assert all(ptrs)
    views = [np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(buf_rows * R,)) for p in ptrs]
    slices = [(lo, min(buf_rows, 150_001 - lo)) for lo in range(0, 150_001, buf_rows)]
    rng = np.random.default_rng(3)
    rng.shuffle(slices)
    try:
        with ctx_for(spec, totals) as ctx:
            tickets = [0] * n_bufs
            for i, (lo, n) in enumerate(slices):
                b = i % n_bufs
                if tickets[b]:
                    t = threading.Thread(target=ctx.host_wait, args=(tickets[b],))     # from another thread, as documented
                    t.start()
                    t.join()
                views[b][:n * R] = rows[lo * R:(lo + n) * R]
                tickets[b] = ctx.scan_host_async(ptrs[b], n, row_base=lo)
                assert tickets[b] == i + 1
            for t in tickets:
                ctx.host_wait(t)
            ctx.host_wait(1)                     # an old ticket is simply done
            with pytest.raises(K.KmdError):
                ctx.host_wait(len(slices) + 5)   # a ticket that was never issued
            assert ctx.counters().rows == 150_001
            check_survivors(ctx.survivors(), want)
    finally:
        for p in ptrs:
            L.kmd_host_free(p)


def test_synthetic_rows_identical_on_host_and_device():
    import torch
    for shape in ("c2", "c5", "b1", "odd"):
        nc, nk, k, b = SHAPES[shape]
        spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=1 << 30, seed=99)
        n, base = 50_003, 123_456_789
        host = spec.rows_host(n, base)
        d = torch.empty(n * spec.row_bytes, dtype=torch.uint8, device="cuda")
        spec.rows_device(d.data_ptr(), n, base)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), host)


@pytest.mark.parametrize("thr", [0.5, 1e-3, 5e-7, 1e-30, 1e-300, 0.0])
def test_prefilter_threshold_sweep(thr):
    """the fp32 screen must never drop a row whatever the prefilter threshold is"""
    spec, rows, totals = make_cohort("c2", 100_000, seed=21, enriched_ppm=20_000)
    want, *_ = oracle_for(spec, totals).diff(rows, 1, 4, thr)
    with ctx_for(spec, totals, thr=thr, max_survivors=1 << 18) as ctx:
        ctx.scan_host(rows)
        check_survivors(ctx.survivors(), want)


def test_threshold_one_keeps_every_row():
    spec, rows, totals = make_cohort("c2", 5_000, seed=2)
    want, *_ = oracle_for(spec, totals).diff(rows, 1, 4, 1.0)
    assert len(want) == 5_000
    with ctx_for(spec, totals, thr=1.0) as ctx:
        ctx.scan_host(rows)
        check_survivors(ctx.survivors(), want)


def adversarial_rows(nc, nk, count_bytes, rng, n):
    """rows built to sit on the decision boundary and at the extremes of the count range.

    Cohort sums beyond the log-factorial table cost the oracle (like the reference,
    log_factorial_table.hpp:14-18) an O(sum) loop of log() per look-up, so rows with large sums are
    kept few: most rows stay below 60 000 per cohort, a dozen reach ~1e6.  Sums never reach 2^31,
    where the reference narrows to `int` (model.hpp:133) and parity is undefined."""
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[count_bytes]
    cmax = int(np.iinfo(dt).max)
    N = nc + nk
    per_elem = max(1, min(cmax, 60_000 // max(nc, nk)))       # keeps a cohort sum <= 60 000
    counts = np.zeros((n, N), dtype=np.uint64)
    kind = rng.integers(0, 6, size=n)
    for i in range(n):
        k = kind[i]
        if k == 0:      # one cohort all zero
            counts[i, nc:] = rng.integers(0, 40, size=nk)
        elif k == 1:
            counts[i, :nc] = rng.integers(0, 40, size=nc)
        elif k == 2:    # large counts, beyond the 10 000-entry table
            counts[i] = rng.integers(per_elem // 2, per_elem, size=N, endpoint=True)
        elif k == 3:    # near the boundary: base b vs b+delta with LR around 12..14
            b = int(rng.integers(1, max(2, min(2000, per_elem - 200))))
            counts[i, :nc] = b
            d = int(np.sqrt(2 * 13.0 * 2 * b * nc / (nc * nk)) * rng.uniform(0.8, 1.2)) + 1
            counts[i, nc:] = min(cmax, b + d)
        elif k == 4:    # single non-zero
            counts[i, rng.integers(0, N)] = rng.integers(1, min(cmax, 500))
        else:           # sparse, a few big values
            j = rng.integers(0, N, size=3)
            counts[i, j] = rng.integers(1, min(cmax, 15_000), size=3, endpoint=True)
    # a dozen rows with cohort sums ~1e6 (4-byte counts only), alternately balanced and one-sided
    if count_bytes == 4:
        for t, i in enumerate(rng.choice(n, size=12, replace=False)):
            counts[i] = 0
            counts[i, :nc] = 1_000_000 // nc
            counts[i, nc:] = (1_000_000 // nk) if t % 2 == 0 else 3
    counts = np.minimum(counts, cmax).astype(dt)
    counts[counts.sum(axis=1) == 0, 0] = 1
    kmers = np.arange(1, n + 1, dtype=np.uint64)
    rows = np.zeros((n, 8 + N * count_bytes), dtype=np.uint8)
    rows[:, :8] = kmers.view(np.uint8).reshape(n, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(n, -1)
    return rows.ravel()


@pytest.mark.parametrize("thr,n", [(0.5, 20_000_000), (1e-3, 20_000_000), (5e-7, 100_000_000), (1e-30, 20_000_000),
                                   (1e-300, 20_000_000)])
def test_screen_never_drops_a_survivor_over_the_whole_count_range(thr, n):
    """up to 1e8 (kc, kk) pairs — one control and one case sample, 4-byte counts, so the two counts ARE the cohort sums — from 1
    to 2^32 - 1 per cohort, most of them placed around the prefilter boundary LR*(thr) at their own magnitude.  The
    scan (fp32 screen -> exact kernel on the candidates) must keep exactly the rows the exact kernel keeps when it is
    run on EVERY row (kmd_model_rows_device): the screen may only err towards keeping a row.  Device vs device, so
    the oracle's O(sum) log-factorial loop is not in the way of sums far beyond the table."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(int(-np.log10(max(thr, 1e-300))) + 5)
    sc, sk = 10**9, 12 * 10**8
    # LR* for this threshold: smallest LR with igamc(0.5, LR) <= thr (bisection on the library's own chi-square tail)
    lo, hi = 0.0, 1.0
    while K._lib.load().kmd_chisqc(1.0, 2 * hi) > thr:
        lo, hi = hi, hi * 2
    for _ in range(100):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if K._lib.load().kmd_chisqc(1.0, 2 * mid) > thr else (lo, mid)
    lr_star = hi
    # total t log-uniform in [2, 2^32), split around the null proportion by a deviation that puts LR near LR*:
    # LR ~ d^2 / (2 t f (1 - f)) for kc = f t + d
    t = torch.exp2(torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 31.0 + 1.0)
    f = sc / (sc + sk)
    d = torch.sqrt(2.0 * lr_star * t * f * (1 - f)) * (0.6 + 0.8 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64))
    d = d * torch.where(torch.rand(n, generator=g, device="cuda") < 0.5, -1.0, 1.0)
    kc = torch.clamp(torch.round(t * f + d), 0, 2**32 - 1)
    kk = torch.clamp(torch.round(t - kc), 0, 2**32 - 1)
    # a tenth of the rows: anything goes (one-sided, tiny, saturated)
    wild = torch.rand(n, generator=g, device="cuda") < 0.1
    kc = torch.where(wild, torch.floor(torch.exp2(torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 32.0)) - 1, kc)
    kk = torch.where(wild, torch.floor(torch.exp2(torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 32.0)) - 1, kk)
    kk = torch.where((kc == 0) & (kk == 0), torch.ones_like(kk), kk)       # kmtricks never emits an all-zero row
    rows = torch.zeros((n, 4), dtype=torch.int32, device="cuda")           # 8-byte k-mer | u32 kc | u32 kk
    rows[:, 0] = torch.arange(n, device="cuda", dtype=torch.int32)
    rows[:, 2] = kc.to(torch.int64).to(torch.int32)                        # wraps to the u32 bit pattern
    rows[:, 3] = kk.to(torch.int64).to(torch.int32)
    del t, d, kc, kk, wild
    p = torch.empty(n, dtype=torch.float64, device="cuda")
    sign = torch.empty(n, dtype=torch.int32, device="cuda")
    mc = torch.empty(n, dtype=torch.float64, device="cuda")
    mk = torch.empty(n, dtype=torch.float64, device="cuda")
    with K.DiffContext(1, 1, sc, sk, kmer_size=31, count_bytes=4, threshold=thr, max_survivors=n) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx._ck(ctx._L.kmd_model_rows_device(ctx._h, rows.data_ptr(), n, p.data_ptr(), sign.data_ptr(), mc.data_ptr(), mk.data_ptr()))
        ctx.scan_device(rows.data_ptr(), n)
        c = ctx.counters()
        want = torch.nonzero(p <= thr).ravel()
        assert c.rows == n and c.survivors == want.numel()
        assert 0.02 * n < c.survivors < 0.98 * n                           # the boundary really runs through the sample
        got = ctx.survivors()
        assert np.array_equal(got["row"], want.cpu().numpy().astype(np.uint64))
        assert np.array_equal(got["p_value"], p[want].cpu().numpy())
        assert c.candidates >= c.survivors


@pytest.mark.parametrize("nc,nk,b", [(10, 10, 4), (3, 17, 2), (8, 8, 1), (100, 100, 4)])
def test_screen_has_no_false_negatives_on_adversarial_rows(nc, nk, b):
    rng = np.random.default_rng(1234 + nc + b)
    n = 20_000
    rows = adversarial_rows(nc, nk, b, rng, n)
    totals = np.full(nc + nk, 10**9, dtype=np.uint64)
    totals[nc:] += 3 * 10**8
    for thr in (5e-7, 1e-2):
        m = oracle.Model(nc, nk, totals)
        want, *_ = m.diff(rows, 1, b, thr)
        with K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=31, count_bytes=b,
                           threshold=thr, max_survivors=1 << 17) as ctx:
            ctx.scan_host(rows)
            got = ctx.survivors()
        assert len(want) > 100
        check_survivors(got, want)


@pytest.mark.parametrize("corr", ["NOTHING", "BONFERRONI", "BENJAMINI", "SIDAK", "HOLM"])
def test_corrections_match_oracle(corr):
    kind = K.CorrectionType[corr]
    spec, rows, totals = make_cohort("c2", 400_000, seed=17, enriched_ppm=3000)
    m = oracle_for(spec, totals)
    want, *_ = m.diff(rows, 1, 4, THR)
    # m chosen so that the correction actually bites for every corrector
    for total in (400_000, 10**9, 10**13):
        sig_level = 0.05 if kind != K.CorrectionType.NOTHING else 1e-12
        keep = oracle.aggregate(want["p"], int(kind), sig_level, total)
        with ctx_for(spec, totals) as ctx:
            ctx.scan_host(rows)
            n_ctl, n_case = ctx.correct(kind, sig_level, total)
            got = ctx.significant()
        exp = want[keep]
        assert len(got) == len(exp)
        assert n_ctl == int((exp["sign"] == 0).sum()) and n_case == int((exp["sign"] != 0).sum())
        # same records in the same order (row order, or ascending p with ties by row)
        assert np.array_equal(got["row"], exp["row"])
        if kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
            assert np.all(np.diff(got["p_value"]) >= 0)


def test_benjamini_breaks_at_first_failure():
    """reference tests/corrector_test.cpp:20-27 semantics on the device path: sorted walk stops at
    the first rejection even if later p-values would pass their own rank test."""
    spec, rows, totals = make_cohort("c2", 300_000, seed=23, enriched_ppm=1000)
    want, *_ = oracle_for(spec, totals).diff(rows, 1, 4, THR)
    ps = np.sort(want["p"])
    total = 300_000
    with ctx_for(spec, totals) as ctx:
        ctx.scan_host(rows)
        sp = ctx.sorted_pvalues()
        assert np.all(np.diff(sp) >= 0) and len(sp) == len(ps)
        n_acc, cut = K.first_reject(sp, K.CorrectionType.BENJAMINI, 0.05, total)
        keep = oracle.aggregate(want["p"], oracle.BENJAMINI, 0.05, total)
        assert n_acc == len(keep)
        a, b = ctx.keep_below(cut)
        assert a + b == n_acc
        assert np.array_equal(ctx.significant()["row"], want["row"][keep])


def test_survivor_overflow_is_reported():
    spec, rows, totals = make_cohort("c2", 50_000, seed=4, enriched_ppm=200_000)
    with ctx_for(spec, totals, max_survivors=64) as ctx:
        ctx.scan_host(rows)
        with pytest.raises(K.KmdError) as e:
            ctx.survivors()
        assert e.value.code == 4


def test_fasta_records_match_reference_format():
    spec, rows, totals = make_cohort("c5", 60_000, seed=8)
    want, *_ = oracle_for(spec, totals).diff(rows, 2, 2, THR)
    with ctx_for(spec, totals) as ctx:
        ctx.scan_host(rows)
        got = ctx.survivors()
    for i in range(min(50, len(got))):
        seq = oracle.kmer_to_string(want["kmer"][i], 63)
        exp = oracle.fasta_record(i, want["p"][i], want["mean_control"][i], want["mean_case"][i], seq)
        rec = K.format_fasta(i, got[i], 63)
        # header prints p with 6 significant digits: identical text unless p sits on a rounding edge
        assert rec.split("\n", 1)[1] == exp.split("\n", 1)[1]
        assert rec.split("_pval=")[0] == exp.split("_pval=")[0]
        assert rec.split("_control=")[1] == exp.split("_control=")[1]


@pytest.mark.parametrize("shape,n", [("c2", 30_000_000), ("c3", 3_000_000)])
def test_large_batch_properties(shape, n):
    """size-independent properties at sizes the oracle would take minutes for: device-generated rows,
    survivors sorted & unique, counters consistent, re-scan idempotent, two half scans == one full."""
    import torch
    nc, nk, k, b = SHAPES[shape]
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=n, seed=1234)
    totals = spec.totals()
    d = torch.empty(n * spec.row_bytes, dtype=torch.uint8, device="cuda")
    spec.rows_device(d.data_ptr(), n)
    torch.cuda.synchronize()
    with ctx_for(spec, totals) as ctx:
        ctx.scan_device(d.data_ptr(), n)
        c = ctx.counters()
        s1 = ctx.survivors()
        assert c.rows == n and c.survivors == len(s1) == c.n_control + c.n_case
        assert c.candidates >= c.survivors
        assert np.all(np.diff(s1["row"].astype(np.int64)) > 0)
        assert np.all(s1["p_value"] <= THR)
        # expected enrichment rate ~ enriched_ppm
        assert 0.5e-3 * n < len(s1) < 2e-3 * n
        # false-positive rate of the screen stays negligible
        assert c.candidates <= 1.05 * c.survivors + 100
        ctx.reset()
        half = (n // 2) // 16 * 16
        ctx.scan_device(d.data_ptr(), half, 0, 0)
        ctx.scan_device(d.data_ptr() + half * spec.row_bytes, n - half, half, 1)
        s2 = ctx.survivors()
        assert np.array_equal(s1["row"], s2["row"]) and np.array_equal(s1["p_value"], s2["p_value"])
        # spot-check 2000 survivors + 2000 random rows against the oracle
        m = oracle_for(spec, totals)
        pick = np.unique(np.concatenate([s1["row"][:: max(1, len(s1) // 2000)],
                                         np.random.default_rng(0).integers(0, n, 2000).astype(np.uint64)]))
        sub = np.concatenate([spec.rows_host(1, int(r)) for r in pick])
        want, *_ = m.diff(sub, spec.kmer_words, spec.count_bytes, THR)
        got_rows = set(s1["row"].tolist())
        assert {int(pick[i]) for i in want["row"]} == {int(r) for r in pick if int(r) in got_rows}
