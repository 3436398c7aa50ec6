"""Multi-GPU plumbing of `kmdiff diff`: one process per GPU, kmtricks partitions sharded across ranks.

The reference's only cross-partition communication is (reference file:line)
  * the sum of per-partition row counts = m of the correctors      include/kmdiff/merge.hpp:316, 390-397
  * the control / case survivor counters                           merge.hpp:319-337
  * one global ascending walk over all survivors for BH / Holm     include/kmdiff/aggregator.hpp:282-310, 325-341
Partitions are disjoint k-mer sets, so there is no data-path collective: each rank scans its own
partitions, then ONE all-gather carries every rank's counters and (for BENJAMINI / HOLM) its survivors'
p-values (<= ~1e-3 of the rows); every rank derives the same m and the same first rejected p-value from
the gathered bytes and keeps its own survivors below it (DeviceExchange).  correct_sharded is the same
decision with the walk on the host (gloo / any backend).

Works on any torch.distributed backend: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist

from .api import CorrectionType, first_reject


def assign_partitions(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Size-balanced partition -> rank assignment (longest-processing-time greedy; ties by index so
    every rank computes the same map).  The reference runs one ThreadPool task per partition in
    index order (merge.hpp:341-388); any disjoint cover gives the same result set."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(sizes[i])
    for r in range(world):
        out[r].sort()
    return out


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def global_counters(rows: int, n_control: int, n_case: int, device, group=None) -> tuple[int, int, int]:
    """all-reduce(sum) of (rows tested, control survivors, case survivors)"""
    if _world(group) == 1:
        return int(rows), int(n_control), int(n_case)
    t = torch.tensor([rows, n_control, n_case], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    r, a, b = t.tolist()
    return int(r), int(a), int(b)


def gather_sorted_pvalues(p_local_sorted: np.ndarray, device, group=None) -> np.ndarray:
    """All ranks' ascending survivor p-values merged into one ascending array (identical on every rank)."""
    world = _world(group)
    p_local_sorted = np.ascontiguousarray(p_local_sorted, dtype=np.float64)
    if world == 1:
        return p_local_sorted
    n_loc = torch.tensor([p_local_sorted.size], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n_loc, group=group)
    sizes_h = sizes.tolist()
    mx = max(1, max(sizes_h))
    pad = torch.full((mx,), float("inf"), dtype=torch.float64, device=device)
    if p_local_sorted.size:
        pad[:p_local_sorted.size] = torch.from_numpy(p_local_sorted).to(device)
    allp = torch.empty(world * mx, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allp, pad, group=group)
    merged = torch.sort(allp).values[:sum(sizes_h)]      # +inf padding sorts last
    return merged.cpu().numpy()


def global_cut(p_local_sorted: np.ndarray, correction: int, significance: float, total_rows: int, device,
               group=None) -> tuple[int, float]:
    """BENJAMINI / HOLM across ranks: (number accepted globally, first rejected p-value).  Every rank
    then keeps its local survivors with p < cut (DiffContext.keep_below)."""
    assert correction in (CorrectionType.BENJAMINI, CorrectionType.HOLM)
    merged = gather_sorted_pvalues(p_local_sorted, device, group)
    return first_reject(merged, correction, significance, total_rows)


class DeviceExchange:
    """do_correction (cmd/diff.hpp:227-260) when partitions are sharded over ranks, without leaving the GPUs and with
    ONE host synchronisation per run (kmd_exchange_* in include/kmdiff_b200.h):

      export   every rank writes its slot {rows, controls, cases, survivors | cap p-values, +inf padded} on the
               context's stream — the survivor count is never read back to size anything;
      gather   one NCCL all-gather of the slots (NVLink), ordered against the context's stream with events;
      correct  m = sum of the rows; BENJAMINI / HOLM: device sort of the union + first rejected rank + p_cut, the other
               correctors: p_cut from m; every rank keeps its survivors below p_cut;
      result   the single read-back: (m, p_cut, kept control / case of this rank, overflow).

    `cap` grows (and the exchange is repeated) when some rank had more survivors than fit a slot; buffers are cached."""

    def __init__(self, device, group=None, cap: int = 1 << 16):
        self.device, self.group = device, group
        self.world = _world(group)
        self.cap = int(cap)
        self.slot = None
        self.all = None
        self.last = None          # the kmd_exchange_decision of the last run
        self.retries = 0          # exchanges repeated because a slot overflowed

    def _buffers(self):
        from ._lib import EXCHANGE_HEADER
        words = EXCHANGE_HEADER + self.cap
        if self.slot is None or self.slot.numel() != words:
            self.slot = torch.zeros(words, dtype=torch.int64, device=self.device)
            self.all = self.slot if self.world == 1 else torch.zeros(words * self.world, dtype=torch.int64, device=self.device)
        return self.slot, self.all

    def correct(self, ctx, correction: int, significance: float):
        """-> (total_rows, (n_control, n_case) kept on this rank); ctx.significant() then returns this rank's records"""
        on_gpu = self.device.type == "cuda"
        while True:
            slot, allb = self._buffers()
            cs = cur = None
            if on_gpu:
                cur = torch.cuda.current_stream(self.device)
                cs = torch.cuda.ExternalStream(ctx.stream_ptr(), device=self.device)
                if cs.cuda_stream == cur.cuda_stream:
                    cs = None
                else:
                    cs.wait_stream(cur)           # the buffers may have just been allocated / zeroed on torch's stream
            ctx.exchange_export(slot.data_ptr(), self.cap)
            if self.world > 1:
                if cs is not None:
                    cur.wait_stream(cs)           # export -> all-gather
                dist.all_gather_into_tensor(allb, slot, group=self.group)
                if cs is not None:
                    cs.wait_stream(cur)           # all-gather -> sort / walk / keep
            ctx.exchange_correct(allb.data_ptr(), self.world, self.cap, correction, significance)
            dec = ctx.exchange_result()
            if not dec.overflow:
                break
            self.retries += 1
            want = int(dec.max_rank_survivors) + int(dec.max_rank_survivors) // 4 + 1
            self.cap = (want + 4095) // 4096 * 4096
        self.last = dec
        return int(dec.total_rows), (int(dec.kept_control), int(dec.kept_case))


def global_gram(ctx, device, group=None, normalize: bool = True) -> tuple[np.ndarray, int]:
    """smartpca's XTX over the sampled k-mers of ALL ranks (--pop-correction with sharded partitions): every rank computes
    the integer co-carrier matrices of its own sampled k-mers per carrier count on the tensor cores
    (kmd_popstrat_gram_partial), one NCCL all-reduce(sum) of those int32 matrices and of the carrier histogram — exact and
    order-free — and the fp64 combination of the sums (kmd_popstrat_gram_combine): bit for bit the matrix one GPU would
    compute from all the k-mers.  The reference accumulates the same sums SNP by SNP in one process
    (thirdparty/hawk/EIG6.0.1-Hawk/src/eigensrc/smartpca.c:907-1025).  -> (xtx, n_used)"""
    if _world(group) == 1:
        return ctx.popstrat_gram(normalize)
    n = ctx.nb_samples
    if n ** 3 * 4 > (2 << 30):
        raise ValueError(f"{n} samples: the dense integer exchange would need {n ** 3 * 4 >> 20} MiB per rank")
    G = torch.empty(n * n * n, dtype=torch.int32, device=device)
    hist = torch.empty(n + 1, dtype=torch.int64, device=device)
    on_gpu = device.type == "cuda"
    if on_gpu:
        torch.cuda.current_stream(device).synchronize()
    ctx.popstrat_gram_partial(G.data_ptr(), hist.data_ptr())      # synchronises the context's stream
    dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    if on_gpu:
        torch.cuda.current_stream(device).synchronize()
    return ctx.popstrat_gram_combine(G.data_ptr(), hist.data_ptr(), normalize)


def correct_sharded(ctx, correction: int, significance: float, device, group=None):
    """do_correction (cmd/diff.hpp:227-260) for one rank's DiffContext when partitions are sharded.
    -> (total_rows, (n_control, n_case) kept on this rank)."""
    c = ctx.counters()
    total_rows, _, _ = global_counters(c.rows, c.n_control, c.n_case, device, group)
    if correction in (CorrectionType.BENJAMINI, CorrectionType.HOLM) and _world(group) > 1:
        _, cut = global_cut(ctx.sorted_pvalues(), correction, significance, total_rows, device, group)
        return total_rows, ctx.keep_below(cut)
    return total_rows, ctx.correct(correction, significance, total_rows)
