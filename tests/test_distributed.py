"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: partition sharding, the global row count
and the cross-rank BENJAMINI / HOLM walk.  Each rank plays one GPU: its "survivors" come from the
oracle on its own partitions; the merged decision must equal the single-process oracle's.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

THR = 0.05 / 100000


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cohort():
    import kmdiff_b200 as K
    spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=240_000, seed=77, enriched_ppm=3000)
    return spec, spec.rows_host(spec.total_rows), spec.totals()


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kmdiff_b200 as K
        import oracle
        from kmdiff_b200 import distributed as D

        spec, rows, totals = _cohort()
        R = spec.row_bytes
        # 7 ragged partitions, sharded by size
        cuts = [0, 10_000, 90_000, 90_001, 130_000, 200_000, 215_000, 240_000]
        sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        mine = D.assign_partitions(sizes, world)[rank]
        m = oracle.Model(10, 10, totals)
        surv, n_rows, n_ctl, n_case = [], 0, 0, 0
        for p in mine:
            s, a, b = m.diff(rows[cuts[p] * R:cuts[p + 1] * R], 1, 4, THR, row_base=cuts[p])
            surv.append(s)
            n_rows += sizes[p]
            n_ctl += a
            n_case += b
        surv = np.concatenate(surv)
        dev = torch.device("cpu")
        total, g_ctl, g_case = D.global_counters(n_rows, n_ctl, n_case, dev)
        res = {"total": total, "ctl": g_ctl, "case": g_case}
        for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
            for m_total in (total, 10**12):
                n_acc, cut = D.global_cut(np.sort(surv["p"]), kind, 0.05, m_total, dev)
                kept = surv["row"][surv["p"] < cut]
                res[f"{int(kind)}_{m_total}"] = (n_acc, cut, kept)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_assign_partitions_is_a_balanced_disjoint_cover():
    from kmdiff_b200 import distributed as D
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        sizes = rng.integers(1, 10**6, size=64).tolist()
        a = D.assign_partitions(sizes, world)
        flat = sorted(i for r in a for i in r)
        assert flat == list(range(64))
        loads = [sum(sizes[i] for i in r) for r in a]
        assert max(loads) - min(loads) <= max(sizes)
        assert a == D.assign_partitions(sizes, world)
    assert D.assign_partitions([5, 5, 5], 8)[3:] == [[] for _ in range(5)]


def test_two_ranks_reach_the_single_process_decision(tmp_path):
    import kmdiff_b200 as K
    import oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npy", allow_pickle=True)[0] for r in range(world)]

    spec, rows, totals = _cohort()
    want, w_ctl, w_case = oracle.Model(10, 10, totals).diff(rows, 1, 4, THR)
    for r in res:
        assert (r["total"], r["ctl"], r["case"]) == (spec.total_rows, w_ctl, w_case)
    for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
        for m_total in (spec.total_rows, 10**12):
            keep = oracle.aggregate(want["p"], int(kind), 0.05, m_total)
            key = f"{int(kind)}_{m_total}"
            assert res[0][key][0] == res[1][key][0] == len(keep)
            assert res[0][key][1] == res[1][key][1]
            got = np.sort(np.concatenate([res[0][key][2], res[1][key][2]]))
            assert np.array_equal(got, np.sort(want["row"][keep]))
    assert 0 < len(oracle.aggregate(want["p"], oracle.BENJAMINI, 0.05, 10**12)) < len(want)


# ---- the same decision on real GPUs over NCCL (needs >= 2 devices; run with `gpurun --gpus N`) ---------

def _nccl_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import kmdiff_b200 as K
        from kmdiff_b200 import distributed as D

        spec, rows, totals = _cohort()
        R = spec.row_bytes
        cuts = [0, 10_000, 90_000, 90_001, 130_000, 200_000, 215_000, 240_000]
        sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        mine = D.assign_partitions(sizes, world)[rank]
        res = {}
        with K.DiffContext(10, 10, int(totals[:10].sum()), int(totals[10:].sum()), threshold=THR, device=rank) as ctx:
            # a small first capacity: the first exchange overflows on some rank and is repeated with a larger one;
            # the context keeps its OWN non-blocking stream for the first two correctors (event ordering against
            # torch's stream and NCCL's), torch's current stream for the rest
            ex = D.DeviceExchange(dev, cap=1024)
            kinds = (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM, K.CorrectionType.BONFERRONI,
                     K.CorrectionType.SIDAK, K.CorrectionType.NOTHING)
            for i, kind in enumerate(kinds):
                if i == 2:
                    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                ctx.reset()
                for p in mine:
                    ctx.scan_host(rows[cuts[p] * R:cuts[p + 1] * R], row_base=cuts[p], partition=p)
                for _ in range(3):       # repeated runs reuse the cached buffers: same decision every time
                    total, kept = ex.correct(ctx, kind, 0.05)
                sig = ctx.significant()
                assert len(sig) == sum(kept)
                res[int(kind)] = (total, kept, sig["row"].copy(), sig["p_value"].copy())
                # the host-side exchange (gather + kmd_first_reject) must reach the same decision
                ctx.reset()
                for p in mine:
                    ctx.scan_host(rows[cuts[p] * R:cuts[p + 1] * R], row_base=cuts[p], partition=p)
                total2, kept2 = D.correct_sharded(ctx, kind, 0.05, dev)
                assert (total2, kept2) == (total, kept)
                assert np.array_equal(ctx.significant()["row"], sig["row"])
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_all_gpus_reach_the_single_process_decision_over_nccl(tmp_path):
    import kmdiff_b200 as K
    import oracle
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = torch.cuda.device_count()
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npy", allow_pickle=True)[0] for r in range(world)]
    spec, rows, totals = _cohort()
    want, _, _ = oracle.Model(10, 10, totals).diff(rows, 1, 4, THR)
    for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM, K.CorrectionType.BONFERRONI,
                 K.CorrectionType.SIDAK, K.CorrectionType.NOTHING):
        keep = oracle.aggregate(want["p"], int(kind), 0.05, spec.total_rows)
        assert all(r[int(kind)][0] == spec.total_rows for r in res)
        got = np.sort(np.concatenate([r[int(kind)][2] for r in res]))
        assert np.array_equal(got, np.sort(want["row"][keep])) and len(keep) > 10
        assert sum(sum(r[int(kind)][1]) for r in res) == len(keep)
        if kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
            for r in res:       # every rank's records come back in ascending p
                assert np.all(np.diff(r[int(kind)][3]) >= 0)


def _gram_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import kmdiff_b200 as K
        from kmdiff_b200 import distributed as D
        spec, rows, totals = _cohort()
        R = spec.row_bytes
        per = (spec.total_rows // world) // 16 * 16
        a, b = rank * per, (spec.total_rows if rank == world - 1 else (rank + 1) * per)
        with K.DiffContext(10, 10, int(totals[:10].sum()), int(totals[10:].sum()), threshold=THR, device=rank) as ctx:
            ctx.popstrat_enable(0.05, seed=9, n_pc=2)
            ctx.scan_host(rows[a * R:b * R], row_base=a)
            xtx, used = D.global_gram(ctx, dev)
        np.save(os.path.join(out_dir, f"gram{rank}.npy"), np.array([{"xtx": xtx, "used": used}], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_global_gram_over_nccl_equals_one_gpu(tmp_path):
    """--pop-correction with partitions sharded over every GPU: the all-reduced integer Gram matrices give every rank the
    XTX one GPU computes from all the rows, bit for bit"""
    import kmdiff_b200 as K
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = torch.cuda.device_count()
    mp.spawn(_gram_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"gram{r}.npy", allow_pickle=True)[0] for r in range(world)]
    spec, rows, totals = _cohort()
    with K.DiffContext(10, 10, int(totals[:10].sum()), int(totals[10:].sum()), threshold=THR) as ctx:
        ctx.popstrat_enable(0.05, seed=9, n_pc=2)
        ctx.scan_host(rows)
        want, used = ctx.popstrat_gram()
    assert used > 1000
    for r in res:
        assert r["used"] == used and np.array_equal(r["xtx"], want)


@pytest.mark.gpu
@pytest.mark.parametrize("own_stream", [True, False])
def test_device_exchange_on_one_gpu_equals_kmd_correct(own_stream):
    """world = 1: the export / sort / walk / keep kernels alone, against kmd_correct and the oracle, for every
    corrector and for m far above the row count (BH / Holm break early); a tiny first capacity forces the
    overflow -> repeat path"""
    import kmdiff_b200 as K
    import oracle
    from kmdiff_b200 import distributed as D
    spec, rows, totals = _cohort()
    want, _, _ = oracle.Model(10, 10, totals).diff(rows, 1, 4, THR)
    dev = torch.device("cuda", 0)
    with K.DiffContext(10, 10, int(totals[:10].sum()), int(totals[10:].sum()), threshold=THR) as ctx:
        if not own_stream:
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ex = D.DeviceExchange(dev, cap=64)
        for kind in K.CorrectionType:
            ctx.reset()
            ctx.scan_host(rows)
            total, kept = ex.correct(ctx, kind, 0.05)
            sig = ctx.significant()
            keep = oracle.aggregate(want["p"], int(kind), 0.05, spec.total_rows)
            assert total == spec.total_rows and sum(kept) == len(keep) == len(sig)
            assert np.array_equal(np.sort(sig["row"]), np.sort(want["row"][keep]))
            assert (ctx.correct(kind, 0.05, total), ) == (kept, )
            assert np.array_equal(ctx.significant()["row"], sig["row"])
            if kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
                assert ex.last.n_accepted == len(keep)
        assert ex.retries == 1 and ex.cap >= len(want)


# ---- DeviceExchange's host logic on CPU: gloo, a stand-in context backed by numpy ------------------------

class _Counters:
    def __init__(self, rows, n_control, n_case, survivors):
        self.rows, self.n_control, self.n_case, self.survivors = rows, n_control, n_case, survivors


class _NumpyCtx:
    """what DeviceExchange needs from a DiffContext — the kmd_exchange_* contract of include/kmdiff_b200.h — restated over
    host arrays (the device kernels have GPU tests); lets gloo carry the slots between CPU ranks"""

    HEADER = 8

    def __init__(self, p, sign, rows):
        self.p, self.sign, self.rows = np.ascontiguousarray(p, dtype=np.float64), sign, rows
        self._dec = None

    def stream_ptr(self):
        return 0

    @staticmethod
    def _view(ptr, words):
        import ctypes
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(words,))

    def exchange_export(self, ptr, cap):
        slot = self._view(ptr, self.HEADER + cap)
        n = len(self.p)
        slot[:self.HEADER] = [self.rows, int((self.sign == 0).sum()), int((self.sign != 0).sum()), n, int(n > cap), 0, 0, 0]
        pv = slot[self.HEADER:].view(np.float64)
        pv[:] = np.inf
        pv[:min(n, cap)] = self.p[:cap]

    def exchange_correct(self, ptr, world, cap, corr, sig):
        import kmdiff_b200 as K
        from kmdiff_b200 import _lib
        allb = self._view(ptr, world * (self.HEADER + cap)).reshape(world, self.HEADER + cap)
        total = int(allb[:, 0].sum())
        dec = _lib.ExchangeDecision()
        dec.total_rows, dec.total_survivors = total, int(allb[:, 3].sum())
        dec.max_rank_survivors, dec.overflow = int(allb[:, 3].max()), int(allb[:, 4].any())
        if corr in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM):
            union = np.sort(allb[:, self.HEADER:].copy().view(np.float64).ravel())
            n_acc, cut = K.first_reject(union, corr, sig, total)
            dec.n_accepted, dec.p_cut = min(n_acc, dec.total_survivors), (cut if n_acc < union.size else np.inf)
        elif corr == K.CorrectionType.BONFERRONI:
            dec.p_cut = sig / total
        elif corr == K.CorrectionType.SIDAK:
            dec.p_cut = 1 - (1 - sig) ** (1.0 / total)
        else:
            dec.p_cut = sig
        k = self.p < dec.p_cut
        dec.kept_control, dec.kept_case = int((k & (self.sign == 0)).sum()), int((k & (self.sign != 0)).sum())
        self._dec = dec

    def exchange_result(self):
        return self._dec


def _exchange_inputs(rank):
    rng = np.random.default_rng(100 + rank)
    n = 5000 - 611 * rank                       # ragged: the shorter ranks are padded with +inf
    p = np.where(rng.random(n) < 0.6, 10.0 ** rng.uniform(-40, -9, n), rng.uniform(1e-7, 5e-7, n))
    p[rng.random(n) < 0.1] = 0.0
    return p, (rng.random(n) < 0.8).astype(np.int32)


def _exchange_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kmdiff_b200 as K
        from kmdiff_b200 import distributed as D
        p, sign = _exchange_inputs(rank)
        ex = D.DeviceExchange(torch.device("cpu"), cap=4096)     # rank 0 holds 5000: the first exchange overflows
        res = {}
        for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM, K.CorrectionType.BONFERRONI):
            for rows in (4_000_000, 10**11):
                res[(int(kind), rows)] = ex.correct(_NumpyCtx(p, sign, rows), kind, 0.05)
        assert ex.retries == 1 and ex.cap >= 5000
        np.save(os.path.join(out_dir, f"ex{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_device_exchange_logic_three_ranks_gloo(tmp_path):
    import kmdiff_b200 as K
    import oracle
    world = 3
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"ex{r}.npy", allow_pickle=True)[0] for r in range(world)]
    ps, signs = zip(*[_exchange_inputs(r) for r in range(world)])
    allp, alls = np.concatenate(ps), np.concatenate(signs)
    for kind in (K.CorrectionType.BENJAMINI, K.CorrectionType.HOLM, K.CorrectionType.BONFERRONI):
        for rows in (4_000_000, 10**11):
            keep = oracle.aggregate(allp, int(kind), 0.05, rows * world)
            want = (int((alls[keep] == 0).sum()), int((alls[keep] != 0).sum()))
            got = [res[r][(int(kind), rows)] for r in range(world)]
            assert all(g[0] == rows * world for g in got)
            assert (sum(g[1][0] for g in got), sum(g[1][1] for g in got)) == want, (kind, rows)
    # the large m makes BH / Holm break early: a strict subset is kept
    keep = oracle.aggregate(allp, oracle.BENJAMINI, 0.05, 3 * 10**11)
    assert 0 < len(keep) < len(allp)
