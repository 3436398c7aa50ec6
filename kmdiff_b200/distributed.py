"""Multi-GPU plumbing of `kmdiff diff`: one process per GPU, kmtricks partitions sharded across ranks.

The reference's only cross-partition communication is (reference file:line)
  * the sum of per-partition row counts = m of the correctors      include/kmdiff/merge.hpp:316, 390-397
  * the control / case survivor counters                           merge.hpp:319-337
  * one global ascending walk over all survivors for BH / Holm     include/kmdiff/aggregator.hpp:282-310, 325-341
Partitions are disjoint k-mer sets, so there is no data-path collective: each rank scans its own
partitions, then (1) one all-reduce(sum) of three u64 counters, and for BENJAMINI / HOLM (2) an
all-gather of the ranks' ascending survivor p-values (<= ~1e-3 of the rows), from which every rank
derives the same first rejected p-value and keeps its own survivors below it.

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
    """BENJAMINI / HOLM across ranks without leaving the GPUs (NCCL): one all-gather of each rank's
    (rows, controls, cases, survivors), one all-gather of the survivors' p-values padded with +inf, one device
    sort of the union, kmd_first_reject_device, then every rank keeps its survivors below the cut.
    Buffers are cached between calls.  The context's stream and torch's current stream are ordered explicitly
    (the context may run on its own non-blocking stream: kmd_set_stream(NULL) does not mean the legacy stream)."""

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.world = _world(group)
        self.meta = torch.zeros(4, dtype=torch.int64, device=device)
        self.meta_all = torch.zeros(4 * self.world, dtype=torch.int64, device=device)
        self.pad = None
        self.allp = None

    def correct(self, ctx, correction: int, significance: float):
        """-> (total_rows, (n_control, n_case) kept on this rank)"""
        c = ctx.counters()
        self.meta.copy_(torch.tensor([c.rows, c.n_control, c.n_case, c.survivors], dtype=torch.int64), non_blocking=True)
        dist.all_gather_into_tensor(self.meta_all, self.meta, group=self.group)
        m = self.meta_all.view(self.world, 4).cpu()
        total_rows = int(m[:, 0].sum())
        if correction not in (CorrectionType.BENJAMINI, CorrectionType.HOLM):
            return total_rows, ctx.correct(correction, significance, total_rows)
        mx = max(1, int(m[:, 3].max()))
        if self.pad is None or self.pad.numel() < mx:
            cap = mx + mx // 4
            self.pad = torch.empty(cap, dtype=torch.float64, device=self.device)
            self.allp = torch.empty(cap * self.world, dtype=torch.float64, device=self.device)
        pad, allp = self.pad[:mx], self.allp[:mx * self.world]
        on_gpu = self.device.type == "cuda"
        pad.fill_(float("inf"))
        if on_gpu:
            torch.cuda.current_stream(self.device).synchronize()      # the fill precedes the context's copy
        ctx.survivor_pvalues_device(pad.data_ptr(), mx)
        if on_gpu:
            ctx.sync()                                                # the copy precedes the all-gather
        dist.all_gather_into_tensor(allp, pad, group=self.group)
        merged = torch.sort(allp).values                      # +inf padding sorts last and never passes a rank test
        if on_gpu:
            torch.cuda.current_stream(self.device).synchronize()      # the sorted union precedes the context's walk
        _, cut = ctx.first_reject_device(merged.data_ptr(), int(m[:, 3].sum()), correction, significance, total_rows)
        return total_rows, ctx.keep_below(cut)


def correct_sharded(ctx, correction: int, significance: float, device, group=None):
    """do_correction (cmd/diff.hpp:227-260) for one rank's DiffContext when partitions are sharded.
    -> (total_rows, (n_control, n_case) kept on this rank)."""
    c = ctx.counters()
    total_rows, _, _ = global_counters(c.rows, c.n_control, c.n_case, device, group)
    if correction in (CorrectionType.BENJAMINI, CorrectionType.HOLM) and _world(group) > 1:
        _, cut = global_cut(ctx.sorted_pvalues(), correction, significance, total_rows, device, group)
        return total_rows, ctx.keep_below(cut)
    return total_rows, ctx.correct(correction, significance, total_rows)
