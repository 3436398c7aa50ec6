"""Pins the oracle's km::KmerMerger restatement (oracle/kmd_merge_oracle.c) and the .kmer file reader:
against the reference's own merge_test.cpp fixture, against vectors made by the unmodified reference
(tests/golden/merge_vectors.json) and, where oracle/_ref exists, against the live reference."""
import json
import os

import numpy as np
import pytest

import oracle
from kmdiff_b200 import kmtricks_io as kio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    with open(os.path.join(GOLDEN, "merge_vectors.json")) as f:
        return json.load(f)["cases"]


def test_reference_fixture_merge_test_cpp(km_out_dir):
    """tests/merge_test.cpp: T == 320 merged rows, totals 160 / 160, nothing significant"""
    FIXTURE = km_out_dir
    cfg = kio.get_kmtricks_config(FIXTURE)
    assert (cfg.kmer_size, cfg.abundance_min, cfg.nb_partitions) == (20, 1, 4)
    tc, tk = kio.get_total_kmer(FIXTURE, 1, 1, cfg.abundance_min)
    assert (tc, tk) == ([160], [160])
    total = 0
    m = oracle.Model(1, 1, np.array(tc + tk, dtype=np.uint64), lf_size=100)
    n_ctl = n_case = 0
    for p in range(cfg.nb_partitions):
        streams = []
        for path in kio.get_partition_paths(FIXTURE, p):
            hdr, rec = kio.read_kmer_file(path)
            assert (hdr.kmer_size, hdr.kmer_slots, hdr.count_slots, hdr.partition) == (20, 1, 4, p)
            streams.append(rec)
        rows = oracle.kmer_merge(streams, 1, 4)
        total += len(rows)
        km = np.ascontiguousarray(rows[:, :8]).view("<u8").ravel()
        assert np.all(km[1:] > km[:-1])                      # ascending, unique
        cnt = np.ascontiguousarray(rows[:, 8:]).view("<u4").reshape(-1, 2)
        assert np.all(cnt.sum(axis=1) > 0)
        # every record of every stream lands in exactly one row
        assert int((cnt[:, 0] > 0).sum()) == len(streams[0]) // 12 and int((cnt[:, 1] > 0).sum()) == len(streams[1]) // 12
        _, c, k = m.diff(rows.ravel(), 1, 4, 0.05 / 10000)
        n_ctl += c
        n_case += k
    assert total == 320
    assert (n_ctl, n_case) == (0, 0)


@pytest.mark.parametrize("ci", range(8))
def test_oracle_matches_reference_vectors(ci):
    c = load_cases()[ci]
    kw = (c["kmer_size"] + 31) // 32
    streams = [np.frombuffer(bytes.fromhex(s), dtype=np.uint8) for s in c["streams"]]
    want = np.frombuffer(bytes.fromhex(c["rows"]), dtype=np.uint8).reshape(c["n_rows"], -1)
    got = oracle.kmer_merge(streams, kw, c["count_bytes"], c["ab_min"])
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("n,k,b,ab", [(2, 20, 4, 1), (9, 31, 2, 2), (33, 31, 1, 1), (5, 63, 4, 3)])
def test_oracle_matches_live_reference(tmp_path, n, k, b, ab):
    if not hasattr(oracle.ref(), "ref_kmer_merge"):
        pytest.skip("oracle/_ref predates ref_kmer_merge")
    rng = np.random.default_rng(n * 1000 + k)
    kw = (k + 31) // 32
    m = 3000
    uni = np.unique(rng.integers(0, 2**min(62, 2 * k), size=m, dtype=np.uint64))
    hi = rng.integers(0, 2**(2 * k - 64), size=len(uni), dtype=np.uint64) if kw == 2 else None
    if kw == 2:
        hi.sort()
    streams, paths = [], []
    for s in range(n):
        sel = rng.random(len(uni)) < 0.3
        rec = np.zeros((int(sel.sum()), 8 * kw + b), dtype=np.uint8)
        rec[:, :8] = uni[sel].view(np.uint8).reshape(-1, 8)
        if kw == 2:
            rec[:, 8:16] = hi[sel].view(np.uint8).reshape(-1, 8)
        cnt = rng.integers(1, min(2**(8 * b) - 1, 9), size=int(sel.sum()), dtype=np.uint64)
        rec[:, 8 * kw:] = cnt.view(np.uint8).reshape(-1, 8)[:, :b]
        if kw == 2:   # order by (hi, lo)
            order = np.lexsort((uni[sel], hi[sel]))
            rec = rec[order]
        streams.append(rec.ravel())
        p = str(tmp_path / f"s{s}.kmer.lz4")
        kio.write_kmer_file(p, rec, k, b, s, 0, lz4=bool(s & 1))
        paths.append(p)
    ab_vec = [ab] * n
    want = oracle.ref_kmer_merge(paths, k, b, ab_vec, cap_rows=len(uni) + 1)
    got = oracle.kmer_merge(streams, kw, b, ab_vec)
    assert np.array_equal(got, want)


def test_rows_to_streams_round_trip():
    """dense rows -> per-sample streams -> merge == the rows that have at least one count"""
    from kmdiff_b200 import SynthSpec
    spec = SynthSpec(4, 5, kmer_size=31, count_bytes=2, total_rows=5000, seed=3)
    rows = spec.rows_host(5000).reshape(-1, spec.row_bytes)
    streams = kio.rows_to_streams(rows, 9, 1, 2)
    got = oracle.kmer_merge(streams, 1, 2)
    assert np.array_equal(got, rows)


@pytest.mark.parametrize("kw,b,n,budget", [(1, 4, 20, 5000), (2, 2, 7, 900), (1, 1, 3, 50), (1, 4, 5, 10**9)])
def test_kmer_range_chunks_cover_the_merge(kw, b, n, budget):
    """iter_kmer_ranges (the host program's feeding rule): merging the ranges one by one gives the rows of the whole
    merge, in order, and no range exceeds the record budget by more than one buffer per stream"""
    rng = np.random.default_rng(kw * 100 + n)
    m = 30_000
    lo = np.unique(rng.integers(0, 2**40, size=m, dtype=np.uint64))
    hi = np.sort(rng.integers(0, 5, size=len(lo), dtype=np.uint64)) if kw == 2 else None
    if kw == 2:
        order = np.lexsort((lo, hi))
        lo, hi = lo[order], hi[order]
    streams = []
    for s in range(n):
        sel = rng.random(len(lo)) < [0.9, 0.05, 0.4][s % 3]
        rec = np.zeros((int(sel.sum()), 8 * kw + b), dtype=np.uint8)
        rec[:, :8] = lo[sel].view(np.uint8).reshape(-1, 8)
        if kw == 2:
            rec[:, 8:16] = hi[sel].view(np.uint8).reshape(-1, 8)
        rec[:, 8 * kw:] = rng.integers(1, 200, size=(int(sel.sum()), 1), dtype=np.uint64).view(np.uint8).reshape(-1, 8)[:, :b]
        streams.append(rec.ravel())
    whole = oracle.kmer_merge(streams, kw, b)
    parts, n_chunks = [], 0
    for chunk in kio.iter_kmer_ranges(streams, kw, b, budget):
        assert sum(c.size for c in chunk) // (8 * kw + b) <= max(budget, n)
        parts.append(oracle.kmer_merge(chunk, kw, b))
        n_chunks += 1
    assert np.array_equal(np.concatenate(parts), whole)
    assert n_chunks > 3 or budget >= 10**9
