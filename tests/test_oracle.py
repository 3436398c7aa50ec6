"""Pins the oracle (oracle/kmd_oracle.c, the CPU restatement used as the parity checker) against
(1) the golden values of the reference's own unit tests, (2) committed vectors produced by the
unmodified reference library (tests/golden/make_golden.py), and (3) — only where oracle/_ref was
built, i.e. where /root/reference exists — the reference library itself on fresh seeded inputs.

CPU only.  Same libm on both sides (x86-64 glibc), no FMA contraction on either: bit-exact.
"""
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(a):
    return np.array([float.fromhex(x) for x in a], dtype=np.float64)


@pytest.fixture(scope="module")
def ref_tests():
    with open(os.path.join(GOLDEN, "reference_tests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def vectors():
    with open(os.path.join(GOLDEN, "ref_vectors.json")) as f:
        return json.load(f)


# ---- (1) the reference's own unit-test goldens --------------------------------------------------

def test_log_factorial_table_goldens(ref_tests):
    """reference tests/factorial_test.cpp:7-16 — EXPECT_EQ on doubles, i.e. bit-exact"""
    g = ref_tests["lf_table"]
    t = oracle.lf_table(g["size"])
    for k, v in g["values"].items():
        assert oracle.lf_lookup(t, int(k)) == v


def test_model_sign_goldens(ref_tests):
    """reference tests/model_test.cpp:45-81"""
    g = ref_tests["model_signs"]
    nb = g["nb"]
    m = oracle.Model(nb, nb, np.full(2 * nb, g["totals"], np.uint64), lf_size=g["preload"])
    for case in g["cases"]:
        row = np.zeros(8 + 8 * nb, np.uint8)
        row[8:] = np.array([case["controls"]] * nb + [case["cases"]] * nb, np.uint32).view(np.uint8)
        _, sign, _, _ = m.rows(row, 1, 4)
        assert sign[0] == {"CONTROL": oracle.CONTROL, "CASE": oracle.CASE, "NO": oracle.NO}[case["sign"]]


def test_known_answers(ref_tests):
    """p-values measured on the reference binary — the reference's tests pin no p-value"""
    for ka in ref_tests["known_answers"]:
        m = oracle.Model(ka["nc"], ka["nk"], np.array(ka["totals_c"] + ka["totals_k"], np.uint64), lf_size=ka["preload"])
        row = np.zeros(8 + 4 * (ka["nc"] + ka["nk"]), np.uint8)
        row[8:] = np.array(ka["controls"] + ka["cases"], np.uint32).view(np.uint8)
        p, s, mc, kk = m.rows(row, 1, 4)
        assert (p[0], s[0], mc[0], kk[0]) == (ka["p"], ka["sign"], ka["mean_control"], ka["mean_case"])


def test_corrector_goldens(ref_tests):
    """reference tests/corrector_test.cpp:9-45"""
    for c in ref_tests["correctors"]:
        oc = oracle.Corrector(oracle.CORRECTIONS[c["type"].lower()], c["threshold"], c["total"])
        for _ in range(c["prime"]):
            oc.apply(0.0)
        assert oc.apply(c["accept"]) and not oc.apply(c["reject"]), c


# ---- (2) committed vectors from the reference library -------------------------------------------

def test_lf_vectors(vectors):
    v = vectors["lf"]
    t = oracle.lf_table(v["size"])
    got = np.array([oracle.lf_lookup(t, i) for i in v["idx"]])
    assert np.array_equal(got, unhex(v["values"]))


def test_chisqc_vectors(vectors):
    v = vectors["chisqc"]
    got = np.array([oracle.chisqc(v["v"], x) for x in unhex(v["x"])])
    assert np.array_equal(got, unhex(v["p"]))


def test_cohort_vectors(vectors):
    for c in vectors["cohorts"]:
        rows = np.frombuffer(bytes.fromhex(c["rows_hex"]), dtype=np.uint8)
        m = oracle.Model(c["nc"], c["nk"], np.array(c["totals"], np.uint64), lf_size=c["lf_size"])
        p, s, mc, kk = m.rows(rows, (c["k"] + 31) // 32, c["count_bytes"])
        assert np.array_equal(p, unhex(c["p"])), c["name"]
        assert s.tolist() == c["sign"], c["name"]
        assert np.array_equal(mc, unhex(c["mean_control"])), c["name"]
        assert np.array_equal(kk, unhex(c["mean_case"])), c["name"]


def test_corrector_vectors(vectors):
    for c in vectors["correctors"]:
        oc = oracle.Corrector(oracle.CORRECTIONS[c["type"]], c["threshold"], c["total"])
        assert [int(oc.apply(float(x))) for x in unhex(c["p"])] == c["apply_in_order"], (c["type"], c["total"])


def test_aggregate_matches_stateful_walk(vectors):
    """kmo_aggregate == the reference's (sorted_)aggregator::run built from the stateful predicates"""
    rng = np.random.default_rng(5)
    p = np.exp(rng.uniform(np.log(1e-14), np.log(1e-4), 500))
    for kind in (oracle.NOTHING, oracle.BONFERRONI, oracle.SIDAK):
        keep = oracle.aggregate(p, kind, 0.05 if kind else 1e-8, 10**6)
        oc = oracle.Corrector(kind, 0.05 if kind else 1e-8, 10**6)
        assert keep.tolist() == [i for i, x in enumerate(p) if oc.apply(float(x))]
    for kind in (oracle.BENJAMINI, oracle.HOLM):
        keep = oracle.aggregate(p, kind, 0.05, 10**6)
        order = np.argsort(p, kind="stable")
        oc = oracle.Corrector(kind, 0.05, 10**6)
        exp = []
        for i in order:
            if not oc.apply(float(p[i])):
                break
            exp.append(int(i))
        assert keep.tolist() == exp and 0 < len(exp) < len(p)


def test_kmer_to_string():
    # A0 C1 T2 G3, first base most significant (kmtricks kmer.hpp:38, 797-810)
    assert oracle.kmer_to_string([0b00011011], 4) == "ACTG"
    assert oracle.kmer_to_string([0], 31) == "A" * 31
    w = [3, 1]  # 128-bit: low word ...G, high word ...C at base index k-1-32
    s = oracle.kmer_to_string(w, 63)
    assert s[-1] == "G" and s[62 - 32] == "C" and s.count("A") == 61


def test_fasta_record_format():
    """aggregator.hpp:55-62: >{n}_pval={:g}_control={size_t}_case={} + 60-column wrap (kseq++)"""
    r = oracle.fasta_record(0, 1.11209e-10, 0.0, 30.0, "A" * 31)
    assert r == ">0_pval=1.11209e-10_control=0_case=30\n" + "A" * 31 + "\n"
    r = oracle.fasta_record(7, 1.0429e-17, 52.9, 37.2, "C" * 63)
    assert r == ">7_pval=1.0429e-17_control=52_case=37.2\n" + "C" * 60 + "\n" + "CCC\n"


# ---- (3) the live reference library, where it was built ------------------------------------------

needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("nc,nk,b,kw", [(10, 10, 4, 1), (100, 100, 4, 1), (50, 50, 2, 2), (7, 6, 1, 1)])
def test_oracle_equals_reference_on_seeded_rows(nc, nk, b, kw):
    rng = np.random.default_rng(nc * 1000 + b)
    n = 20_000
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[b]
    lam = rng.gamma(2.0, 3.0, size=(n, 1))
    counts = rng.poisson(lam, size=(n, nc + nk))
    boost = rng.random(n) < 0.05
    counts[boost, nc:] += rng.poisson(30.0, size=(int(boost.sum()), nk))
    counts = np.minimum(counts, np.iinfo(dt).max).astype(dt)
    rows = np.zeros((n, 8 * kw + (nc + nk) * b), np.uint8)
    rows[:, 8 * kw:] = counts.view(np.uint8).reshape(n, -1)
    totals = rng.integers(9 * 10**8, 11 * 10**8, nc + nk).astype(np.uint64)
    want = oracle.ref_model_rows(nc, nk, totals, 10000, rows, kw, b, threads=4)
    got = oracle.Model(nc, nk, totals).rows(rows, kw, b)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@needs_ref
def test_oracle_chisqc_equals_reference_on_dense_grid():
    xs = np.exp(np.linspace(np.log(1e-8), np.log(1500.0), 4000))
    assert all(oracle.chisqc(1.0, float(x)) == oracle.ref_chisqc(1.0, float(x)) for x in xs)


# ---- the reference BINARIES (4-byte and 2-byte count builds) against the oracle, on the CPU ---------------

@pytest.mark.skipif(not os.path.exists(oracle.REF_BIN), reason="oracle/_ref/kmdiff not built (no /root/reference)")
@pytest.mark.parametrize("nc,nk,k,b", [(10, 10, 31, 4), (6, 9, 31, 2), (5, 5, 63, 2)])
def test_reference_binary_spill_files_match_oracle(tmp_path, nc, nk, k, b):
    """`kmdiff diff --keep-tmp` of the unmodified reference: the prefilter survivors of its spill files (k-mer, p,
    sign, means) equal the oracle's bit for bit.  k = 63: the reference writes only the low 64 bits of the k-mer
    (SURVEY.md §8 hazard 1), which is what is compared."""
    import subprocess

    import kmdiff_b200 as K
    from kmdiff_b200 import build as kb
    from kmdiff_b200 import kmtricks_io as kio

    binary = oracle.REF_BIN if b == 4 else oracle.REF_BIN16
    if not os.path.exists(binary):
        pytest.skip(f"{binary} not built")
    n_rows = 60_000
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=n_rows, seed=17, enriched_ppm=4000)
    d, out = str(tmp_path / "run"), str(tmp_path / "out")
    kio.write_run_dir(d, spec, n_rows, 1)
    r = subprocess.run([binary, "diff", "--km-run", d, "-1", str(nc), "-2", str(nk), "-o", out, "--keep-tmp", "-t", "1",
                        "-s", "0.05", "-u", "100000"], capture_output=True, text=True, timeout=600)
    assert "[error]" not in r.stderr, r.stderr[-600:]
    rows = spec.rows_host(n_rows)
    want, n_ctl, n_case = oracle.Model(nc, nk, spec.totals()).diff(rows, spec.kmer_words, b, 0.05 / 100000)
    assert f"{len(want)}/{n_rows} significant k-mers." in r.stderr
    assert f"Before correction: {n_ctl} (control), {n_case} (case)." in r.stderr
    raw = subprocess.run([kb.build_cli(), "dev-cat", os.path.join(out, "partitions", "p0_uncorrected"), "0", "lz4"],
                         capture_output=True, timeout=60).stdout
    # the reference dumps 8 k-mer bytes whatever k is; built WITH_POPSTRAT every record also carries `u16 N | N x f64
    # counts` (kmer.hpp:118-123)
    n = nc + nk
    rec = 8 + 28 + 2 + 8 * n
    assert len(raw) == rec * len(want) and len(want) > 30
    a = np.frombuffer(raw, dtype=np.uint8).reshape(-1, rec)
    assert np.array_equal(np.ascontiguousarray(a[:, :8]).view("<u8").ravel(), want["kmer"][:, 0])
    assert np.array_equal(np.ascontiguousarray(a[:, 8:16]).view("<f8").ravel(), want["p"])
    assert np.array_equal(np.ascontiguousarray(a[:, 16:20]).view("<i4").ravel(), want["sign"])
    assert np.array_equal(np.ascontiguousarray(a[:, 20:28]).view("<f8").ravel(), want["mean_control"])
    assert np.array_equal(np.ascontiguousarray(a[:, 28:36]).view("<f8").ravel(), want["mean_case"])
    assert np.all(np.ascontiguousarray(a[:, 36:38]).view("<u2").ravel() == n)
    dt = {2: "<u2", 4: "<u4"}[b]
    counts = np.ascontiguousarray(rows.reshape(n_rows, -1)[:, 8 * spec.kmer_words:]).view(dt).reshape(n_rows, n)
    assert np.array_equal(np.ascontiguousarray(a[:, 38:]).view("<f8").reshape(-1, n), counts[want["row"]].astype(np.float64))


@pytest.mark.skipif(not os.path.exists(oracle.REF_BIN), reason="oracle/_ref/kmdiff not built (no /root/reference)")
@pytest.mark.parametrize("corr", ["bonferroni", "benjamini", "sidak", "holm", "disabled"])
def test_reference_binary_fasta_matches_oracle(tmp_path, corr):
    """The FASTA files of the unmodified reference for every corrector against oracle.aggregate + oracle.fasta_record:
    same records (header text after the running index, sequence); BENJAMINI / HOLM additionally in the same order
    (ascending p).  m is made large (10^7-row matrix header trick is not available, so --cutoff 1 and a tight -s make
    the correctors bite)."""
    import subprocess

    import kmdiff_b200 as K
    from kmdiff_b200 import kmtricks_io as kio

    nc = nk = 10
    n_rows = 80_000
    spec = K.SynthSpec(nc, nk, kmer_size=31, count_bytes=4, total_rows=n_rows, seed=23, enriched_ppm=3000)
    d, out = str(tmp_path / "run"), str(tmp_path / "out")
    kio.write_run_dir(d, spec, n_rows, 3)
    # significance levels at which each corrector rejects a part of the prefilter survivors (Sidak's
    # 1 - (1 - t)^(1/m) underflows to 0 for t = 1e-12)
    sig, cutoff = {"bonferroni": "1e-12", "benjamini": "1e-12", "holm": "1e-12", "sidak": "1e-3", "disabled": "1e-9"}[corr], "1"
    r = subprocess.run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out, "-c", corr, "-t", "1",
                        "-s", sig, "-u", cutoff], capture_output=True, text=True, timeout=600)
    assert "[error]" not in r.stderr, r.stderr[-600:]
    rows = spec.rows_host(n_rows)
    want, _, _ = oracle.Model(nc, nk, spec.totals()).diff(rows, 1, 4, float(sig) / float(cutoff))
    keep = oracle.aggregate(want["p"], oracle.CORRECTIONS[corr], float(sig), n_rows)
    assert 0 < len(keep) < len(want) or corr in ("disabled",)          # the corrector rejects part of the survivors
    exp = {"control": [], "case": []}
    for rcd in want[keep]:
        rec = oracle.fasta_record(0, rcd["p"], rcd["mean_control"], rcd["mean_case"], oracle.kmer_to_string(rcd["kmer"][:1], 31))
        head, body = rec.split("\n", 1)
        exp["control" if rcd["sign"] == oracle.CONTROL else "case"].append((head.split("_", 1)[1], body.replace("\n", "")))
    for name in ("control", "case"):
        got, head = [], None
        with open(os.path.join(out, f"{name}_kmers.fasta")) as f:
            for line in f:
                line = line.rstrip("\n")
                if line.startswith(">"):
                    head = line.split("_", 1)[1]
                else:
                    got.append((head, line))
        if corr in ("benjamini", "holm"):
            assert [g[1] for g in got] == [e[1] for e in exp[name]] or sorted(got) == sorted(exp[name])
            assert sorted(got) == sorted(exp[name])
        else:
            assert sorted(got) == sorted(exp[name])
    assert f"Significant k-mers: {len(exp['control'])} (control), {len(exp['case'])} (case)." in r.stderr
