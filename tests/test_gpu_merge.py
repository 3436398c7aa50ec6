"""GPU parity of the per-sample count-file input: the device merge tree behind
kmd_merge_rows_host / kmd_merge_scan_host against the oracle's km::KmerMerger restatement, the
reference's golden vectors and its merge_test.cpp fixture.  Everything here is integer work: bit-exact."""
import json
import os

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from kmdiff_b200 import kmtricks_io as kio
from conftest import SHAPES, make_cohort, rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
THR = 0.05 / 100000


def ctx_for(nc, nk, k, b, totals=None, thr=THR, **kw):
    if totals is None:
        totals = np.full(nc + nk, 10**9, dtype=np.uint64)
    return K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=k, count_bytes=b,
                         threshold=thr, **kw)


def test_reference_fixture_merge_test_cpp(km_out_dir):
    """tests/merge_test.cpp through the CUDA path: 320 rows, 0 control / 0 case survivors"""
    FIXTURE = km_out_dir
    cfg = kio.get_kmtricks_config(FIXTURE)
    tc, tk = kio.get_total_kmer(FIXTURE, 1, 1, cfg.abundance_min)
    with K.DiffContext(1, 1, tc[0], tk[0], kmer_size=cfg.kmer_size, count_bytes=4, threshold=0.05 / 10000, lf_size=100) as ctx:
        base = 0
        for p in range(cfg.nb_partitions):
            streams = [kio.read_kmer_file(q)[1] for q in kio.get_partition_paths(FIXTURE, p)]
            want = oracle.kmer_merge(streams, 1, 4)
            assert np.array_equal(ctx.merge_rows_host(streams), want)
            base += ctx.merge_scan_host(streams, row_base=base, partition=p)
        c = ctx.counters()
    assert (c.rows, c.n_control, c.n_case) == (320, 0, 0)


@pytest.mark.parametrize("ci", range(8))
def test_merge_matches_reference_vectors(ci):
    with open(os.path.join(GOLDEN, "merge_vectors.json")) as f:
        c = json.load(f)["cases"][ci]
    streams = [np.frombuffer(bytes.fromhex(s), dtype=np.uint8) for s in c["streams"]]
    want = np.frombuffer(bytes.fromhex(c["rows"]), dtype=np.uint8).reshape(c["n_rows"], -1)
    n = c["n_samples"]
    with ctx_for(n // 2, n - n // 2, c["kmer_size"], c["count_bytes"]) as ctx:
        got = ctx.merge_rows_host(streams, c["ab_min"])
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("shape,n_rows", [("c2", 200_003), ("c5", 50_021), ("b1", 120_000), ("odd", 70_001),
                                          ("tiny", 9_000), ("b2odd", 33_333), ("c3", 6_007), ("c4", 3_001)])
def test_merge_inverts_rows_to_streams(shape, n_rows):
    """dense synthetic rows -> per-sample streams -> device merge == the rows; then the merged rows are scanned
    on the device and the survivors equal the oracle's on the original matrix"""
    spec, rows, totals = make_cohort(shape, n_rows, seed=23)
    n = spec.nb_controls + spec.nb_cases
    streams = kio.rows_to_streams(rows, n, spec.kmer_words, spec.count_bytes)
    dense = rows.reshape(-1, spec.row_bytes)
    with ctx_for(spec.nb_controls, spec.nb_cases, spec.kmer_size, spec.count_bytes, totals) as ctx:
        got = ctx.merge_rows_host(streams)
        assert np.array_equal(got, dense)
        assert np.array_equal(got, oracle.kmer_merge(streams, spec.kmer_words, spec.count_bytes))
        assert ctx.merge_scan_host(streams, row_base=5, partition=3) == n_rows
        c = ctx.counters()
        surv = ctx.survivors()
    want, w_ctl, w_case = oracle.Model(spec.nb_controls, spec.nb_cases, totals).diff(rows, spec.kmer_words, spec.count_bytes, THR)
    assert (c.rows, c.survivors, c.n_control, c.n_case) == (n_rows, len(want), w_ctl, w_case)
    assert np.array_equal(surv["row"], want["row"] + 5)
    assert np.array_equal(surv["kmer"], want["kmer"])
    assert np.array_equal(surv["sign"], want["sign"])
    assert np.all(surv["partition"] == 3)
    assert rel_err(surv["p_value"], want["p"]).max() <= 1e-9


@pytest.mark.parametrize("n,k,b", [(2, 31, 4), (3, 20, 1), (37, 31, 2), (64, 63, 4), (129, 31, 4), (400, 31, 1)])
def test_merge_ab_min_and_ragged_streams(n, k, b):
    """abundance thresholds per sample, empty streams, streams of very different lengths, sample counts that are
    not powers of two (lone runs in the merge tree)"""
    rng = np.random.default_rng(n * 7 + k)
    kw = (k + 31) // 32
    m = 20_000
    lo = np.unique(rng.integers(0, 2**min(62, 2 * k), size=m, dtype=np.uint64))
    if kw == 2:
        hi = np.sort(rng.integers(0, 4, size=len(lo), dtype=np.uint64))   # many equal high words
        order = np.lexsort((lo, hi))
        lo, hi = lo[order], hi[order]
    streams = []
    for s in range(n):
        dens = [0.0, 0.9, 0.02, 0.3][s % 4] if n > 3 else 0.5
        sel = rng.random(len(lo)) < dens
        rec = np.zeros((int(sel.sum()), 8 * kw + b), dtype=np.uint8)
        rec[:, :8] = lo[sel].view(np.uint8).reshape(-1, 8)
        if kw == 2:
            rec[:, 8:16] = hi[sel].view(np.uint8).reshape(-1, 8)
        cnt = rng.integers(1, min(2**(8 * b) - 1, 12), size=int(sel.sum()), dtype=np.uint64)
        rec[:, 8 * kw:] = cnt.view(np.uint8).reshape(-1, 8)[:, :b]
        streams.append(rec.ravel())
    ab = rng.integers(1, 6, size=n).astype(np.uint32)
    want = oracle.kmer_merge(streams, kw, b, ab)
    with ctx_for(n // 2, n - n // 2, k, b) as ctx:
        got = ctx.merge_rows_host(streams, ab)
        assert got.shape == want.shape and np.array_equal(got, want)
        # a second call on the same context reuses the device buffers
        got1 = ctx.merge_rows_host(streams[::-1], ab[::-1].copy())
        assert np.array_equal(got1, oracle.kmer_merge(streams[::-1], kw, b, ab[::-1].copy()))


def test_merge_empty_inputs():
    with ctx_for(2, 2, 31, 4) as ctx:
        empty = [np.empty(0, np.uint8)] * 4
        assert ctx.merge_rows_host(empty).shape[0] == 0
        assert ctx.merge_scan_host(empty) == 0
        assert ctx.counters().rows == 0
        one = [np.empty(0, np.uint8)] * 3 + [np.frombuffer((7).to_bytes(8, "little") + (9).to_bytes(4, "little"), dtype=np.uint8)]
        rows = ctx.merge_rows_host(one)
        assert rows.shape == (1, 24) and rows[0, :8].view("<u8")[0] == 7 and list(rows[0, 8:].view("<u4")) == [0, 0, 0, 9]
        # every count below the threshold: the k-mer has no solid sample and is dropped (recurrence_min = 1)
        assert ctx.merge_rows_host(one, ab_min=10).shape[0] == 0


def test_chunked_merge_equals_whole_partition():
    """the host program feeds a partition in k-mer ranges: merging [.., bound] and (bound, ..] separately gives
    the same rows as one call"""
    spec, rows, totals = make_cohort("c2", 100_000, seed=4)
    n = 20
    streams = kio.rows_to_streams(rows, n, 1, 4)
    dense = rows.reshape(-1, spec.row_bytes)
    bound = np.ascontiguousarray(dense[41_234, :8]).view("<u8")[0]
    first, second = [], []
    for s in streams:
        rec = s.reshape(-1, 12)
        km = np.ascontiguousarray(rec[:, :8]).view("<u8").ravel()
        cut = int(np.searchsorted(km, bound, side="right"))
        first.append(rec[:cut].ravel())
        second.append(rec[cut:].ravel())
    with ctx_for(10, 10, 31, 4, totals) as ctx:
        a = ctx.merge_rows_host(first)
        b = ctx.merge_rows_host(second)
    assert len(a) == 41_235
    assert np.array_equal(np.concatenate([a, b]), dense)


@pytest.mark.parametrize("frames", [False, True])
def test_prefetched_chunks_merge_like_plain_calls(frames):
    """kmd_merge_prefetch: the next chunk's H2D copies run on the copy stream while the current chunk is merged; a merge
    call given the prefetched pointers uses that copy, any other call ignores it — results identical either way"""
    import ctypes
    import pyarrow as pa
    from kmdiff_b200 import _lib
    L = _lib.load()
    spec, rows, totals = make_cohort("c2", 90_000, seed=4)
    R = spec.row_bytes
    want, *_ = oracle.Model(10, 10, totals).diff(rows, 1, 4, THR)
    cuts = [0, 20_000, 20_001, 55_000, 90_000]
    chunks, keep = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        streams = kio.rows_to_streams(rows[a * R:b * R], 20, 1, 4)
        if frames:
            streams = [np.frombuffer(pa.compress(s.tobytes(), codec="lz4", asbytes=True), dtype=np.uint8) for s in streams]
        # pinned copies, like the host program's buffers
        pinned = []
        for s_ in streams:
            ptr = L.kmd_host_alloc(max(1, s_.size))
            view = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(1, s_.size),))[:s_.size]
            view[:] = s_
            pinned.append(view)
            keep.append(ptr)
        chunks.append(pinned)
    merge = (lambda ctx, st, base: ctx.merge_scan_lz4_host(st, row_base=base)[0]) if frames else \
            (lambda ctx, st, base: ctx.merge_scan_host(st, row_base=base))
    try:
        with ctx_for(10, 10, 31, 4, totals) as ctx:
            base = 0
            ctx.merge_prefetch(chunks[0], frames=frames)
            for i, st in enumerate(chunks):
                if i + 1 < len(chunks):
                    ctx.merge_prefetch(chunks[i + 1], frames=frames)
                if i == 2:
                    # a call with OTHER buffers in between must not pick the waiting prefetch up ...
                    other = [np.array(s_) for s_ in st]
                    assert merge(ctx, other, base) == cuts[i + 1] - cuts[i]
                    ctx.reset()
                    base = 0
                    # ... and everything is redone from the start without prefetching: the stale slots are simply reused
                    for j, st2 in enumerate(chunks[:i]):
                        base += merge(ctx, [np.array(s_) for s_ in st2], base)
                base += merge(ctx, st, base)
            assert ctx.counters().rows == 90_000
            got = ctx.survivors()
        assert np.array_equal(got["row"], want["row"]) and np.array_equal(got["kmer"], want["kmer"])
    finally:
        for ptr in keep:
            L.kmd_host_free(ptr)


# ---- the two merge algorithms (kmd_union.cuh / kmd_merge.cuh) -------------------------------------------------

def _random_streams(rng, n, k, b, m, dens, max_count=12):
    """n sorted record streams over a universe of m random k-mers; dens(s) = share of the universe sample s holds"""
    kw = (k + 31) // 32
    lo = np.unique(rng.integers(0, 2**min(62, 2 * k), size=m, dtype=np.uint64))
    hi = None
    if kw == 2:
        hi = np.sort(rng.integers(0, 3, size=len(lo), dtype=np.uint64))
        order = np.lexsort((lo, hi))
        lo, hi = lo[order], hi[order]
    streams = []
    for s in range(n):
        sel = rng.random(len(lo)) < dens(s)
        rec = np.zeros((int(sel.sum()), 8 * kw + b), dtype=np.uint8)
        rec[:, :8] = lo[sel].view(np.uint8).reshape(-1, 8)
        if kw == 2:
            rec[:, 8:16] = hi[sel].view(np.uint8).reshape(-1, 8)
        cnt = rng.integers(1, min(2**(8 * b) - 1, max_count), size=int(sel.sum()), dtype=np.uint64)
        rec[:, 8 * kw:] = cnt.view(np.uint8).reshape(-1, 8)[:, :b]
        streams.append(rec.ravel())
    return streams, kw


@pytest.mark.parametrize("n,k,b,m,kind", [
    (20, 31, 4, 300_000, "dense"),       # every sample holds ~90 % of the k-mers: runs hardly grow from level to level
    (20, 31, 4, 300_000, "sparse"),      # ~1/n each: the union is nearly the concatenation
    (200, 31, 4, 40_000, "mixed"),       # c3 width, densities 0 .. 0.9, empty streams
    (7, 63, 2, 120_000, "dense"),        # two key words, 2-byte counts, lone runs on two levels
    (5, 63, 4, 80_000, "dense"),         # two key words inside 20-byte records (the first level reads them there)
    (400, 31, 4, 5_000, "dense"),        # c4 width: 1608-byte rows, 32 of them per assembly tile
    (33, 20, 1, 90_000, "mixed"),        # 1-byte counts (records of 9 bytes: nothing is aligned)
    (2, 31, 4, 50_000, "one"),           # one stream with records: its keys are the rows
    (2, 32, 4, 70_000, "dense"),         # k = 32: all 64 key bits in use
])
@pytest.mark.parametrize("rows_mode", ["assembled", "scattered"])
def test_union_tree_equals_pairwise_tree_and_oracle(n, k, b, m, kind, rows_mode, monkeypatch):
    """every record solid -> the union tree (k-mers merged with duplicates dropped, counts scattered); the pairwise
    (k-mer, payload) tree on the same streams and the oracle's km::KmerMerger give the same rows, byte for byte"""
    rng = np.random.default_rng(n * 1000 + k + b)
    dens = {"dense": lambda s: 0.9, "sparse": lambda s: 1.0 / n, "mixed": lambda s: [0.0, 0.9, 0.02, 0.3, 0.6][s % 5],
            "one": lambda s: float(s == 1)}[kind]
    streams, kw = _random_streams(rng, n, k, b, m, dens)
    want = oracle.kmer_merge(streams, kw, b)
    # dense cohorts get their rows assembled in shared memory, sparse ones scattered into zeroed rows: force each
    monkeypatch.setenv("KMD_UNION_ROWS", {"assembled": "1", "scattered": "2"}[rows_mode])
    with ctx_for(n // 2, n - n // 2, k, b) as ctx:
        got = ctx.merge_rows_host(streams)
        info = ctx.merge_info()
        assert info["algo"] == "union" and info["rows"] == rows_mode
        assert got.shape == want.shape and np.array_equal(got, want)
        n_rec = sum(s.size for s in streams) // (8 * kw + b)
        if n > 1:
            # level 0 reads every record's key once; the root run is the row set
            assert info["keys_in"] >= n_rec and info["keys_out"] >= len(want)
        # the same context again (buffers, look-back words and tickets are reused), streams in another order
        got1 = ctx.merge_rows_host(streams[::-1])
        assert np.array_equal(got1, oracle.kmer_merge(streams[::-1], kw, b))
        assert ctx.counters().rows == 0      # also: no lookup failure flagged by the scatter
    monkeypatch.setenv("KMD_MERGE_PAIRWISE", "1")
    with ctx_for(n // 2, n - n // 2, k, b) as ctx:
        got2 = ctx.merge_rows_host(streams)
        assert ctx.merge_info()["algo"] == "pairwise"
    assert np.array_equal(got2, want)


def test_union_tree_falls_back_when_a_record_is_not_solid():
    """a count below its sample's abundance_min is absent as far as the row's existence goes (merge.hpp:204-223,
    239-240): the union tree notices and the pairwise tree, which carries the solid flags, takes over"""
    rng = np.random.default_rng(5)
    streams, kw = _random_streams(rng, 6, 31, 4, 50_000, lambda s: 0.5, max_count=6)
    ab = np.array([1, 1, 3, 1, 1, 2], dtype=np.uint32)
    want = oracle.kmer_merge(streams, kw, 4, ab)
    with ctx_for(3, 3, 31, 4) as ctx:
        got = ctx.merge_rows_host(streams, ab)
        assert ctx.merge_info()["algo"] == "pairwise"
        assert np.array_equal(got, want)
        # thresholds every count meets -> the union tree again on the same context
        got = ctx.merge_rows_host(streams, 1)
        assert ctx.merge_info()["algo"] == "union"
        assert np.array_equal(got, oracle.kmer_merge(streams, kw, 4))
        # ab_min = 0 and a zero count: solid (0 >= 0), the row exists with a zero in it
        z = [s.copy() for s in streams]
        z[0].reshape(-1, 12)[::7, 8:] = 0
        zab = np.array([0, 1, 1, 1, 1, 1], dtype=np.uint32)
        got = ctx.merge_rows_host(z, zab)
        assert np.array_equal(got, oracle.kmer_merge(z, kw, 4, zab))


def test_union_tree_picks_the_row_writer_by_density(monkeypatch):
    rng = np.random.default_rng(9)
    for kind, dens, want_mode in (("dense", lambda s: 0.8, "assembled"), ("sparse", lambda s: 0.05, "scattered")):
        streams, kw = _random_streams(rng, 16, 31, 4, 60_000, dens)
        want = oracle.kmer_merge(streams, kw, 4)
        with ctx_for(8, 8, 31, 4) as ctx:
            got = ctx.merge_rows_host(streams)
            assert ctx.merge_info()["rows"] == want_mode, kind
        assert np.array_equal(got, want)
        # 4-byte counts: the first level reads the k-mers inside the records; with the knob off they are unpacked first
        monkeypatch.setenv("KMD_UNION_RAW_LEAF", "0")
        with ctx_for(8, 8, 31, 4) as ctx:
            assert np.array_equal(ctx.merge_rows_host(streams), want)
            assert ctx.merge_info()["algo"] == "union"
        monkeypatch.delenv("KMD_UNION_RAW_LEAF")


def test_merge_scan_device_equals_merge_scan_host():
    """records already in HBM (kmd_merge_scan_device): same rows tested, same survivors as from host streams"""
    import torch
    spec, rows, totals = make_cohort("c2", 150_000, seed=11)
    streams = kio.rows_to_streams(rows, 20, 1, 4)
    want, w_ctl, w_case = oracle.Model(10, 10, totals).diff(rows, 1, 4, THR)
    d = torch.from_numpy(np.concatenate(streams)).cuda()
    cnt = [s.size // 12 for s in streams]
    with ctx_for(10, 10, 31, 4, totals) as ctx:
        assert ctx.merge_scan_device(d.data_ptr(), cnt, row_base=7, partition=2) == 150_000
        c = ctx.counters()
        surv = ctx.survivors()
        assert (c.rows, c.survivors, c.n_control, c.n_case) == (150_000, len(want), w_ctl, w_case)
        assert np.array_equal(surv["row"], want["row"] + 7) and np.array_equal(surv["kmer"], want["kmer"])
        ctx.reset()
        assert ctx.merge_scan_host(streams, row_base=7, partition=2) == 150_000
        surv2 = ctx.survivors()
    assert np.array_equal(surv2["row"], surv["row"]) and np.array_equal(surv2["p_value"], surv["p_value"])
