#!/usr/bin/env python
"""bench.py — k-mer rows tested per second (Poisson LRT + correction) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU `kmdiff diff`
    torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU (NCCL)

One "step" = one pass of the hot path over one batch of synthetic kmtricks-format rows held by each
GPU: reset -> LRT scan of the batch (screen + exact fp64 test of the candidates + prefilter) ->
global row count (NCCL all-reduce when N > 1) -> significance correction on the device (sort /
compaction) -> counts of significant control / case k-mers back on the host.

Workload at N=1: BASELINE.json configs[1] ("synthetic 10 control + 10 case, k=31, 1e9 k-mer rows,
Poisson LRT + Bonferroni"), 88 B/row = 88 GB resident in HBM.  Every rank holds its own 1e9-row
partition set (weak scaling; partitions are disjoint, no data-path collective).

`value`  : rows/s with the rows already resident in HBM (device generated, seeded).
`e2e`    : rows/s through the C-ABI with HOST buffers (kmd_scan_host from pinned memory, survivors
           and significant k-mers read back), H2D/D2H inside the timed region, on a smaller batch.
`roofline`: lrt_scan kernel, algorithmic bytes R*rows / its CUDA-event duration vs measured HBM peak.
`cpu_baseline`: the reference's own `kmdiff diff` binary (oracle/_ref/kmdiff) on the host cores,
           on a bounded prefix of the same synthetic matrix written as a kmtricks run dir.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (nb_controls, nb_cases, k, count_bytes, rows, correction)
    "c2": (10, 10, 31, 4, 1_000_000_000, "bonferroni"),
    "c3": (100, 100, 31, 4, 150_000_000, "benjamini"),   # 2e9 rows = 1.6 TB do not fit one GPU
    "c5": (50, 50, 63, 2, 500_000_000, "bonferroni"),    # per-GPU share of 4e9 rows over 8 GPUs
    # BASELINE.json configs[3]: population stratification, 10 PCs, 1e6 sampled k-mers, logistic GLM
    "c4": (200, 200, 31, 4, 60_000_000, "bonferroni"),
    # the c2 cohort as `kmdiff count` leaves it: per-sample sorted (k-mer, count) files, no matrices
    # (BASELINE.json configs[0]'s input layout; km::KmerMerger row of SURVEY.md §8a) — host streams in
    "c2m": (10, 10, 31, 4, 8_000_000, "bonferroni"),
}
C4_NPC = 10
C4_SAMPLED = 1_000_000
SIGNIFICANCE = 0.05
CUTOFF = 100000


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the unmodified `kmdiff diff` binary on the host cores
# ------------------------------------------------------------------------------------------------------

def reference_available():
    import oracle
    return os.path.exists(oracle.REF_BIN)


def make_reference_run_dir(cfg_name: str, sample_rows: int, partitions: int) -> tuple[str, int]:
    """kmtricks run dir holding rows [0, sample_rows) of the workload's synthetic matrix"""
    import kmdiff_b200 as K
    from kmdiff_b200 import kmtricks_io as kio
    nc, nk, k, b, rows, _ = CONFIGS[cfg_name]
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=rows, seed=42)
    d = tempfile.mkdtemp(prefix="kmdiff_ref_run_", dir=os.environ.get("KMD_TMP", "/tmp"))
    kio.write_run_dir(d, spec, sample_rows, partitions, count_files=cfg_name.endswith("m"))
    return d, spec.row_bytes


def run_reference_once(run_dir: str, cfg_name: str, threads: int, pop_args=()) -> float:
    import oracle
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    binary = oracle.REF_BIN if b == 4 else oracle.REF_BIN16
    out = tempfile.mkdtemp(prefix="kmdiff_ref_out_", dir=os.environ.get("KMD_TMP", "/tmp"))
    opt = os.path.join(run_dir, "kmdiff-count.opt")
    if os.path.exists(opt):
        os.remove(opt)   # reference quirk: matrix mode truncates it and the next run then fails
    cmd = [binary, "diff", "--km-run", run_dir, "-1", str(nc), "-2", str(nk), "-o", out, "-t", str(threads),
           "-c", corr, "-s", str(SIGNIFICANCE), "-u", str(CUTOFF)] + list(pop_args)
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    shutil.rmtree(out, ignore_errors=True)
    if r.returncode != 0 or "[error]" in r.stderr:
        raise RuntimeError(f"reference kmdiff failed: {r.stderr[-500:]}")
    return dt


def reference_sample_rows(cfg_name: str, cores: int) -> int:
    nc, nk, k, b, rows, _ = CONFIGS[cfg_name]
    # ~3 M rows/s/core at N=20, scaling down with the row width; aim at ~5 s per run
    per_core = int(3_000_000 * 20 / (nc + nk))
    return int(min(rows, 32_000_000, max(200_000, per_core * min(cores, 64) * 2)))


def bench_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    base = {"metric": "kmer_rows_tested_per_sec", "unit": "rows/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    import oracle
    binary = oracle.REF_BIN if b == 4 else oracle.REF_BIN16
    if not os.path.exists(binary):
        print(json.dumps({"impl": "reference", "unavailable": f"{binary} not built (make -C oracle refbin)"}))
        return
    cores = os.cpu_count() or 1
    sample = reference_sample_rows(cfg_name, cores)
    parts = max(cores, 8)
    run_dir, R = make_reference_run_dir(cfg_name, sample, parts)
    try:
        for _ in range(args.warmup):
            run_reference_once(run_dir, cfg_name, cores)
        t = [run_reference_once(run_dir, cfg_name, cores) for _ in range(args.steps)]
    finally:
        shutil.rmtree(run_dir, ignore_errors=True)
    total = sum(t)
    value = sample * args.steps / total
    line = dict(base)
    line.update({
        "value": value, "ms_per_step": 1e3 * total / args.steps,
        "config": {"workload": f"{cfg_name}: synthetic {nc}+{nk} samples, k={k}, {b}-byte counts, {corr}",
                   "rows_per_step": sample, "partitions": parts, "row_bytes": R,
                   "note": "reference `kmdiff diff` binary (file read + LRT + correction + FASTA) on a "
                           "bounded prefix of the workload's synthetic matrix, " +
                           ("uncompressed per-sample count files merged by km::KmerMerger" if cfg_name.endswith("m")
                            else "uncompressed matrices")},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "reference",
                         "sample": f"{sample} rows in {parts} partitions per step"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm, per-sample count-file input (km::KmerMerger row): host streams -> device merge -> scan -> correction
# ------------------------------------------------------------------------------------------------------

def bench_merge(args) -> None:
    import numpy as np
    import torch

    import kmdiff_b200 as K
    from kmdiff_b200 import kmtricks_io as kio

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py: c2m is a single-GPU bench line")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — kmdiff_b200 has no CPU fallback")
    cfg_name = args.config
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    if args.rows:
        rows = args.rows
    n = nc + nk
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=rows, seed=42)
    totals = spec.totals()
    kw, rec = spec.kmer_words, 8 * spec.kmer_words + b
    # the host program's reader state, prepared once: every sample's records of a k-mer range per merge call
    # The streams live in pinned host memory, like the reader buffers of `kmdiff-b200 diff`.
    import ctypes
    from kmdiff_b200 import _lib
    L = _lib.load()
    chunk_rows = max(1, min(rows, (24 << 20) // n))
    chunks, n_records, pinned = [], 0, []
    for lo in range(0, rows, chunk_rows):
        dense = spec.rows_host(min(chunk_rows, rows - lo), lo)
        streams = kio.rows_to_streams(dense, n, kw, b)
        n_records += sum(s.size for s in streams) // rec
        total = sum(s.size for s in streams)
        hptr = L.kmd_host_alloc(max(total, 1))
        if not hptr:
            raise SystemExit("kmd_host_alloc failed")
        pinned.append(hptr)
        hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(total, 1),))
        views, at = [], 0
        for s_ in streams:
            hbuf[at:at + s_.size] = s_
            views.append(hbuf[at:at + s_.size])
            at += s_.size
        chunks.append(views)
    thr = SIGNIFICANCE / CUTOFF
    ctx = K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=k, count_bytes=b, threshold=thr)
    corr_kind = K.correction_from_str(corr)
    stage_ms = {"unpack": 0.0, "passes": 0.0, "rows": 0.0, "scan": 0.0}
    n_passes = [0]

    def step(timed=False):
        ctx.reset()
        base = 0
        for streams in chunks:
            base += ctx.merge_scan_host(streams, row_base=base)
            if timed:
                m = ctx.merge_ms()
                for key in ("unpack", "passes", "rows"):
                    stage_ms[key] += m[key]
                n_passes[0] = m["n_passes"]
                ctx.sync()
                stage_ms["scan"] += ctx.scan_kernel_ms()
        c = ctx.counters()
        res = ctx.correct(corr_kind, SIGNIFICANCE, c.rows)
        sig = ctx.significant()
        return res, c, sig

    for _ in range(max(3, args.warmup)):
        res, c, sig = step()
    assert c.rows == rows, (c.rows, rows)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, c, sig = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches0) // args.steps
    step(timed=True)                      # per-stage device times, outside the timed region (adds syncs)
    value = rows * args.steps / dt
    peak, peak_src = load_peaks()
    pass_bytes = n_passes[0] * n_records * 2 * (8 * kw + 8)     # every pass reads and writes (key, payload)
    achieved = pass_bytes / (stage_ms["passes"] * 1e-3) / 1e9 if stage_ms["passes"] else None

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            if reference_available():
                sample = min(rows, reference_sample_rows("c2", cores) // 4)
                parts = max(cores, 8)
                run_dir, _ = make_reference_run_dir(cfg_name, sample, parts)
                try:
                    run_reference_once(run_dir, cfg_name, cores)
                    dtc = run_reference_once(run_dir, cfg_name, cores)
                finally:
                    shutil.rmtree(run_dir, ignore_errors=True)
                cpu = {"value": sample / dtc, "unit": "rows/s", "cores": cores, "kind": "reference",
                       "sample": f"`kmdiff diff -t {cores}` (oracle/_ref/kmdiff, km::KmerMerger path) on the first {sample} "
                                 f"rows as {parts} partitions x {n} uncompressed .kmer files, {dtc:.2f} s"}
        except Exception as e:
            cpu = {"value": None, "unit": "rows/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}

    line = {
        "metric": "kmer_rows_tested_per_sec", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg_name}: synthetic {nc}+{nk} samples, k={k}, {b}-byte counts as per-sample sorted count "
                               f"streams (no matrices), device N-way merge + Poisson LRT + {corr}",
                   "rows_per_gpu": rows, "records_per_gpu": n_records, "record_bytes": rec, "row_bytes": spec.row_bytes,
                   "merge_calls_per_step": len(chunks), "merge_passes": n_passes[0],
                   "l2": f"every merge call works on {n_records // max(1, len(chunks)) * 32 >> 20} MiB of (key, payload) arrays, above the 126 MB L2",
                   "survivors_per_gpu": int(c.survivors), "significant": [int(res[0]), int(res[1])]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None, "traffic": None, "kernel": "mrg::merge_pass_kernel",
                     "ms_per_launch": stage_ms["passes"] / max(1, n_passes[0] * len(chunks)),
                     "algorithmic_bytes_per_launch": pass_bytes // max(1, n_passes[0] * len(chunks)), "peak_source": peak_src,
                     "stage_ms_per_step": stage_ms},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": n_records * rec, "d2h_bytes_per_step": int(sig.nbytes + 48),
                "rows_per_step": rows, "steps": args.steps,
                "api": "kmd_merge_scan_host (host record streams) + kmd_correct + kmd_significant; this path has no "
                       "device-resident variant, value == e2e"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    ctx.close()
    for hptr in pinned:
        L.kmd_host_free(hptr)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def bench_b200(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import kmdiff_b200 as K
    from kmdiff_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — kmdiff_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout: keep fd 1 on stderr until the one JSON line is due
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    cfg_name = args.config
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    if args.rows:
        rows = args.rows
    corr_kind = K.correction_from_str(corr)
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=rows * world, seed=42)
    R = spec.row_bytes
    totals = spec.totals()
    thr = SIGNIFICANCE / CUTOFF

    stream = torch.cuda.current_stream()
    ctx = K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=k, count_bytes=b,
                        threshold=thr, device=local_rank, max_survivors=max(1 << 22, rows // 200))
    ctx.set_stream(stream.cuda_stream)
    popstrat = cfg_name == "c4"
    pop_ms = {"gram_kernel": [], "gram_call": [], "pca_host": [], "glm": [], "sampled": 0, "used": 0}
    if popstrat:
        if world > 1:
            raise SystemExit("bench.py: c4 is a single-GPU bench line (the multi-GPU Gram sum is covered by tests)")
        ctx.popstrat_enable(C4_SAMPLED / rows, seed=42, is_diploid=True, n_pc=C4_NPC, max_sampled=2 * C4_SAMPLED)

    # ---- resident workload: this rank's rows, generated on the device (seeded, counter-based) ----
    data = torch.empty(rows * R, dtype=torch.uint8, device=dev)
    spec.rows_device(data.data_ptr(), rows, row_base=rank * rows, device=local_rank, stream=stream.cuda_stream)
    torch.cuda.synchronize()

    exchange = None
    if world > 1:
        from kmdiff_b200.distributed import DeviceExchange
        exchange = DeviceExchange(dev)

    def pop_stage(timed: bool):
        # do_pop (cmd/diff.hpp:168-224): Gram matrix of the sampled k-mers -> PCs -> null model -> GLM per survivor
        t0 = time.perf_counter()
        xtx, used = ctx.popstrat_gram()
        t1 = time.perf_counter()
        _, Z = K.smartpca_pcs(xtx, 10)
        t2 = time.perf_counter()
        ctx.popstrat_init(Z, totals)
        ctx.sync()
        t3 = time.perf_counter()
        ctx.popstrat_apply()
        if timed:
            ctx.sync()      # wall clock around a synchronised launch: the IRLS kernel is the only work in it
            t4 = time.perf_counter()
            pop_ms["gram_kernel"].append(ctx.popstrat_gram_ms())
            pop_ms["gram_call"].append(1e3 * (t1 - t0))
            pop_ms["pca_host"].append(1e3 * (t2 - t1))
            pop_ms["glm"].append(1e3 * (t4 - t3))
            pop_ms["used"] = int(used)

    last_scan_ms = [0.0]

    def step():
        ctx.reset()
        ctx.scan_device(data.data_ptr(), rows, row_base=rank * rows, partition=rank)
        c = ctx.counters()
        total_rows = c.rows
        last_scan_ms[0] = ctx.scan_kernel_ms()     # before the Gram kernel reuses the context's event pair
        if popstrat:
            pop_stage(True)
        if world > 1:
            # m of the correctors = rows of all ranks; BH/Holm: one ascending walk over all ranks' survivors,
            # exchanged and ranked on the devices (kmdiff_b200.distributed.DeviceExchange)
            _, kept = exchange.correct(ctx, corr_kind, SIGNIFICANCE)
            return kept, c
        return ctx.correct(corr_kind, SIGNIFICANCE, total_rows), c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        res, c = step()
    launches0 = ctx.launch_count()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        res, c = step()
        scan_ms.append(last_scan_ms[0])
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = (ctx.launch_count() - launches0) // args.steps
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = rows * world * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers ----
    e2e_rows = min(rows, args.e2e_rows, (8 << 30) // R)      # at most 8 GiB of pinned host rows
    L = _lib.load()
    hptr = L.kmd_host_alloc(e2e_rows * R)
    if not hptr:
        raise SystemExit("kmd_host_alloc failed")
    import ctypes
    hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint8)), shape=(e2e_rows * R,))
    torch.cuda.synchronize()
    hbuf[:] = data[:e2e_rows * R].cpu().numpy()      # same rows as the resident workload's prefix
    d2h = 0

    def e2e_step():
        nonlocal d2h
        ctx.reset()
        ctx.scan_host(hptr, e2e_rows, row_base=rank * rows, partition=rank)
        cc = ctx.counters()
        total = cc.rows * world
        if popstrat:
            pop_stage(False)
        if world > 1:
            _, r = exchange.correct(ctx, corr_kind, SIGNIFICANCE)      # same global decision as the resident step
        else:
            r = ctx.correct(corr_kind, SIGNIFICANCE, total)
        sig = ctx.significant()
        d2h = sig.nbytes + 48 * 3
        return r, len(sig)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_res, n_sig = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_rows * world * args.e2e_steps / float(te.item())
    L.kmd_host_free(hptr)

    # ---- roofline of the dominant kernel (lrt_scan) ----
    peak, peak_src = load_peaks()
    scan_avg_ms = sum(scan_ms) / len(scan_ms)
    achieved = rows * R / (scan_avg_ms * 1e-3) / 1e9
    traffic = None
    traffic_src = None
    import glob
    for prof in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_lrt_scan_ncu.json")), reverse=True):
        try:
            with open(prof) as f:
                pj = json.load(f)
            if pj.get("config") == cfg_name and pj.get("rows"):
                # ncu --set full capture of the same kernel at pj["rows"] rows; DRAM bytes scale with rows
                traffic = pj["dram_bytes_per_launch"] * (rows / pj["rows"])
                traffic_src = os.path.relpath(prof, ROOT)
                break
        except Exception:
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the reference binary on a bounded prefix of the same matrix (N=1 only) ----
    cpu = None
    import oracle as _oracle_paths                      # paths of the reference binaries only
    have_smartpca = os.path.exists(_oracle_paths.REF_SMARTPCA)
    if world == 1 and not args.no_cpu_baseline and (not popstrat or have_smartpca):
        cores = os.cpu_count() or 1
        try:
            if reference_available():
                sample = reference_sample_rows(cfg_name, cores)
                pop_args, what = (), "`kmdiff diff"
                if popstrat:
                    # the reference's whole --pop-correction pipeline (Sampler, HAWK smartpca, evec2pca.perl, GLM) with
                    # the workload's sampling rate, on at most ~1 GB of rows
                    sample = int(min(sample, (1 << 30) // R))
                    pop_args = ("--pop-correction", "--kmer-pca", repr(C4_SAMPLED / rows), "--n-pc", str(C4_NPC), "--random-seed", "42")
                    what = "`kmdiff diff --pop-correction` (+ oracle/_ref/smartpca)"
                parts = max(cores, 8)
                run_dir, _ = make_reference_run_dir(cfg_name, sample, parts)
                try:
                    run_reference_once(run_dir, cfg_name, cores, pop_args)           # warm the page cache
                    dt = run_reference_once(run_dir, cfg_name, cores, pop_args)
                finally:
                    shutil.rmtree(run_dir, ignore_errors=True)
                cpu = {"value": sample / dt, "unit": "rows/s", "cores": cores, "kind": "reference",
                       "sample": f"{what} -t {cores}` (oracle/_ref/kmdiff) on the first {sample} rows of the "
                                 f"workload, {parts} uncompressed partitions, {dt:.2f} s"}
            else:
                import oracle
                sample = min(rows, 2_000_000)
                hrows = data[:sample * R].cpu().numpy()
                m = oracle.Model(nc, nk, totals)
                t0 = time.perf_counter()
                m.diff(hrows, spec.kmer_words, b, thr)
                dt = time.perf_counter() - t0
                cpu = {"value": sample / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                       "sample": f"oracle/kmd_oracle.c on the first {sample} rows, in memory, {dt:.2f} s"}
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "rows/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}

    line = {
        "metric": "kmer_rows_tested_per_sec", "value": value, "unit": "rows/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{cfg_name}: synthetic {nc}+{nk} samples, k={k}, {b}-byte counts, Poisson LRT + {corr}",
            "rows_per_gpu": rows, "row_bytes": R, "bytes_per_gpu": rows * R, "significance": SIGNIFICANCE,
            "cutoff": CUTOFF, "partition_sharding": f"{world} ranks x {rows} rows, disjoint",
            "l2": "inputs (>= 13 GB per step) far exceed the 126 MB L2; no flush needed",
            "survivors_per_gpu": int(c.survivors), "significant": [int(res[0]), int(res[1])],
            "screen_candidates": int(c.candidates),
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "lrt_scan_kernel", "ms_per_launch": scan_avg_ms,
                     "algorithmic_bytes_per_launch": rows * R, "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": e2e_rows * R,
                "d2h_bytes_per_step": int(d2h), "rows_per_step": e2e_rows, "steps": args.e2e_steps,
                "api": ("kmd_scan_host + kmd_popstrat_{gram,init,apply} + kmd_correct + kmd_significant" if popstrat else
                        "kmd_scan_host + kmd_correct + kmd_significant") + " (pinned host buffer)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if popstrat:
        n = nc + nk
        k = args.steps
        avg = lambda v: sum(v[-k:]) / max(1, len(v[-k:]))
        gram_ms = avg(pop_ms["gram_kernel"])
        used = pop_ms["used"]
        ops = 2.0 * n * n * used                       # algorithmic int8 multiply-adds x2 of X X^T
        bf16 = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                bf16 = float(json.load(f)["bf16_tflops"])
        except Exception:
            pass
        int8_peak = 2.0 * bf16 if bf16 else 4500.0
        line["popstrat"] = {
            "n_pc": C4_NPC, "sampled_kmers_used": used, "survivors_fitted": int(c.survivors),
            "gram": {"bound": "tensor", "kernel": "grm_umma_kernel", "ms_per_launch": gram_ms,
                     "achieved": ops / (gram_ms * 1e-3) / 1e12 if gram_ms else None, "peak": int8_peak, "unit": "TOP/s",
                     "frac": (ops / (gram_ms * 1e-3) / 1e12) / int8_peak if gram_ms else None,
                     "peak_source": "2 x measured dense bf16 (MEASURED_PEAKS.json): int8 tensor rate" if bf16 else "nominal 4.5 POP/s",
                     "call_ms": avg(pop_ms["gram_call"]),
                     "note": "2*N^2*m ops; N=400 occupies 4 of 4 M tiles x 512 TMEM columns after padding to 512"},
            "pca_host_ms": avg(pop_ms["pca_host"]),
            "glm": {"kernel": "irls_kernel", "ms": avg(pop_ms["glm"]),
                    "kmers_per_s": int(c.survivors) / (avg(pop_ms["glm"]) * 1e-3) if pop_ms["glm"] else None},
        }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the config's)")
    ap.add_argument("--e2e-rows", type=int, default=1 << 25)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    elif args.config.endswith("m"):
        bench_merge(args)
    else:
        bench_b200(args)


if __name__ == "__main__":
    main()
