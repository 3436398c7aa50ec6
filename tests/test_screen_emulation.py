"""The fp32 prefilter screen of lrt_scan_kernel (kmdiff_b200/csrc/kmd_scan.cuh: screen_pass) replayed in numpy float32
on the CPU, against the likelihood ratio in float64.

The kernel keeps a row unless  lr2 + margin < lr2_cut  with
    lr2    = kc*log2(kc/ec) + kk*log2(kk/ek)            (fp32, lg2.approx + div.approx)
    margin = 0.08 + 2.5e-6 * (|kc*lc| + |kk*lk| + kc + kk)
and lr2_cut a hair below LR*(thr)/ln 2.  A true survivor has LR >= LR*, so the screen is safe iff the fp32 value is
never more than `margin` below the exact one.  The two approximate instructions are not reproducible off the GPU; they
enter with their documented error bounds (CUDA C Programming Guide, intrinsic functions: __fdividef 2 ulp,
__log2f 2^-22 absolute on [0.5, 2] and 2 ulp elsewhere), charged in the adverse direction on top of the correctly
rounded numpy float32 evaluation.  The GPU tests (test_gpu_parity.py) attack the same bound with the real instructions.
"""
import numpy as np
import pytest

F = np.float32
ULP = 2.0 ** -23


def screen_terms(kc64, kk64, frac_c, frac_k):
    """screen_pass's arithmetic, every operation rounded to float32 -> (lr2, margin, lc, lk) as float64 views"""
    with np.errstate(divide="ignore", invalid="ignore"):
        kc, kk = kc64.astype(F), kk64.astype(F)
        t = kc + kk
        ec, ek = t * F(frac_c), t * F(frac_k)
        lc = np.where(kc > 0, np.log2(kc / ec, dtype=F), F(0)).astype(F)
        lk = np.where(kk > 0, np.log2(kk / ek, dtype=F), F(0)).astype(F)
        tc, tk = kc * lc, kk * lk
        lr2 = tc + tk
        margin = F(0.08) + F(2.5e-6) * (np.abs(tc) + np.abs(tk) + t)
    return lr2.astype(np.float64), margin.astype(np.float64), lc.astype(np.float64), lk.astype(np.float64)


def exact_lr2(kc, kk, sc, sk):
    kc, kk = kc.astype(np.float64), kk.astype(np.float64)
    t = kc + kk
    ec, ek = t * (sc / (sc + sk)), t * (sk / (sc + sk))
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(kc > 0, kc * np.log2(kc / ec), 0.0)
        b = np.where(kk > 0, kk * np.log2(kk / ek), 0.0)
    return a + b


def sample_pairs(rng, n, sc, sk, lr_star):
    """(kc, kk) up to 2^32 - 1 per cohort: 80 % around the boundary LR = lr_star at their own magnitude, 20 % wild"""
    t = np.exp2(rng.uniform(1.0, 32.0, n))
    f = sc / (sc + sk)
    d = np.sqrt(2.0 * lr_star * t * f * (1 - f)) * rng.uniform(0.5, 1.5, n) * rng.choice([-1.0, 1.0], n)
    kc = np.clip(np.round(t * f + d), 0, 2**32 - 1)
    kk = np.clip(np.round(t - kc), 0, 2**32 - 1)
    wild = rng.random(n) < 0.2
    kc = np.where(wild, np.floor(np.exp2(rng.uniform(0, 32, n))) - 1, kc)
    kk = np.where(wild, np.floor(np.exp2(rng.uniform(0, 32, n))) - 1, kk)
    kk = np.where((kc == 0) & (kk == 0), 1, kk)
    return kc.astype(np.uint64), kk.astype(np.uint64)


@pytest.mark.parametrize("sc,sk", [(10**10, 10**10), (10**9, 12 * 10**8), (3 * 10**9, 10**8), (10**6, 10**11)])
@pytest.mark.parametrize("lr_star", [0.23, 12.6, 70.0, 690.0])       # LR*(thr) for thr = 0.5, 5e-7, 1e-30, 1e-300
def test_fp32_screen_stays_within_its_margin(sc, sk, lr_star):
    rng = np.random.default_rng(int(lr_star * 10) + sc % 97)
    kc, kk = sample_pairs(rng, 2_000_000, float(sc), float(sk), lr_star)
    frac_c, frac_k = F(sc / (sc + sk)), F(sk / (sc + sk))
    lr2, margin, lc, lk = screen_terms(kc, kk, frac_c, frac_k)
    want = exact_lr2(kc, kk, float(sc), float(sk))
    kcf, kkf = kc.astype(np.float64), kk.astype(np.float64)
    # adverse error of the two approximate instructions, per cohort, in log2 units:
    #   div.approx: 2 ulp relative on the ratio -> 2*ULP/ln 2 absolute in log2
    #   lg2.approx: 2^-22 absolute near 1, 2 ulp relative elsewhere
    approx = lambda l: 2 * ULP / np.log(2) + np.maximum(2.0 ** -22, 2 * ULP * np.abs(l))
    hw = kcf * approx(lc) + kkf * approx(lk)
    shortfall = (want - lr2) + hw          # how far below the exact value the kernel's lr2 can be
    assert np.all(np.isfinite(lr2)) and np.all(np.isfinite(margin))
    worst = np.max(shortfall / margin)
    assert worst < 1.0, worst              # the margin covers it everywhere ...
    assert worst < 0.75                    # ... with room to spare
    # the decision itself: every pair at or above the boundary passes the screen (lr2_cut <= LR*/ln 2)
    lr2_cut = lr_star / np.log(2)
    survivors = want >= lr2_cut
    assert survivors.sum() > 100_000
    assert np.all((lr2 - hw + margin)[survivors] >= lr2_cut)
    # and it is a screen, not a pass-through: far below the boundary rows are dropped
    far = want < 0.5 * lr2_cut - 1.0
    if far.any():
        assert np.mean((lr2 + margin)[far] < lr2_cut) > 0.5


def test_screen_edge_cases():
    sc, sk = 1e9, 1.2e9
    kc = np.array([0, 5, 1, 0, 2**32 - 1, 2**32 - 1, 1], dtype=np.uint64)
    kk = np.array([7, 0, 1, 2**32 - 1, 0, 2**32 - 1, 2**32 - 1], dtype=np.uint64)
    lr2, margin, _, _ = screen_terms(kc, kk, F(sc / (sc + sk)), F(sk / (sc + sk)))
    want = exact_lr2(kc, kk, sc, sk)
    assert np.all(np.isfinite(lr2)) and np.all(np.abs(want - lr2) < margin)
