#!/usr/bin/env python
"""bench.py — k-mer rows tested per second (Poisson LRT + correction) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU `kmdiff diff`
    torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU (NCCL)

One "step" = one pass of the hot path over one batch of synthetic kmtricks-format rows held by each
GPU: reset -> LRT scan of the batch (screen + exact fp64 test of the candidates + prefilter) ->
device-side exchange (kmdiff_b200.distributed.DeviceExchange: one all-gather of every rank's row /
survivor counters and survivor p-values, device sort of the union, first rejected rank, keep) -> ONE
host synchronisation that returns m, the cut and the counts of significant control / case k-mers.

Headline workload at every N: BASELINE.json configs[2], the shape the north-star target is quoted on
("synthetic 100 + 100 samples, k=31, ..., Benjamini-Hochberg correction"), 808 B/row, 1.5e8 rows =
121 GB resident per GPU (2e9 rows = 1.6 TB do not fit 8 x 180 GB; every rank holds its own disjoint
partition set: weak scaling, no data-path collective).  The line also carries, under "c2", the same
measurement of configs[1] ("10 control + 10 case, k=31, 1e9 k-mer rows, Poisson LRT + Bonferroni",
88 GB resident per GPU), the headline of round 1, and at N = 1 under "c2m" the count-file input of
that cohort (`--config c2m`, `--config c3m` alone: the km::KmerMerger row — per-sample sorted record
streams -> device N-way merge -> the same scan; `value` from records resident in HBM through
kmd_merge_scan_device, `e2e` from pinned host streams through kmd_merge_scan_host).

`value`   : rows/s with the rows already resident in HBM (device generated, seeded).
`e2e`     : rows/s through the C-ABI with HOST buffers (kmd_scan_host from pinned memory, exchange,
            significant records read back), H2D/D2H inside the timed region, on a smaller batch.
`e2e_cli` : the drop-in program itself, file -> FASTA: wall time of `kmdiff-b200 diff` against the
            reference binary `kmdiff diff -t $(nproc)` on the same run directory (N = 1 only).
`roofline`: lrt_scan kernel, algorithmic bytes R*rows / its CUDA-event duration vs measured HBM peak.
`cpu_baseline`: the reference's own `kmdiff diff` binary (oracle/_ref/kmdiff) on the host cores, on a
            bounded prefix of the same synthetic matrix written as a kmtricks run dir; its FASTA k-mer
            set is compared with ours on the same prefix (`parity_on_prefix`).
`decision_check`: outside the timed region the same survivors are corrected again with the walk on
            the host (correct_sharded: gather + kmd_first_reject); m, the kept counts and the kept
            rows must be identical or the bench fails.
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
    # name: (nb_controls, nb_cases, k, count_bytes, rows per GPU, correction)
    "c2": (10, 10, 31, 4, 1_000_000_000, "bonferroni"),
    "c3": (100, 100, 31, 4, 150_000_000, "benjamini"),   # 2e9 rows = 1.6 TB do not fit 8 x 180 GB
    "c5": (50, 50, 63, 2, 500_000_000, "bonferroni"),    # per-GPU share of 4e9 rows over 8 GPUs
    # BASELINE.json configs[3]: population stratification, 10 PCs, 1e6 sampled k-mers, logistic GLM
    "c4": (200, 200, 31, 4, 60_000_000, "bonferroni"),
    # the c2 cohort as `kmdiff count` leaves it: per-sample sorted (k-mer, count) files, no matrices
    # (BASELINE.json configs[0]'s input layout; km::KmerMerger row of SURVEY.md §8a) — host streams in
    "c2m": (10, 10, 31, 4, 8_000_000, "bonferroni"),
    # the same for the north-star cohort width: 200 streams, 8 levels of the merge tree
    "c3m": (100, 100, 31, 4, 1_000_000, "benjamini"),
}
HEADLINE = "c3"
SECONDARY = "c2"
C4_NPC = 10
C4_SAMPLED = 1_000_000
SIGNIFICANCE = 0.05
CUTOFF = 100000
SEED = 42
ENRICHED_PPM = 1000
CLI = os.path.join(ROOT, "kmdiff_b200", "kmdiff-b200")


def workload_name(cfg_name: str) -> str:
    nc, nk, k, b, _, corr = CONFIGS[cfg_name]
    if cfg_name.endswith("m"):
        return (f"{cfg_name}: synthetic {nc}+{nk} samples, k={k}, {b}-byte counts as per-sample sorted count streams "
                f"(no matrices), N-way merge + Poisson LRT + {corr}")
    return f"{cfg_name}: synthetic {nc}+{nk} samples, k={k}, {b}-byte counts, Poisson LRT + {corr}"


def row_bytes(cfg_name: str) -> int:
    nc, nk, k, b, _, _ = CONFIGS[cfg_name]
    return 8 * ((k + 31) // 32) + (nc + nk) * b


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def scratch_dir() -> str:
    """where run directories go: $KMD_TMP, else /dev/shm when it has room (the GPU boxes' / is a slow virtual disk
    and both arms read the same files), else /tmp"""
    env = os.environ.get("KMD_TMP")
    if env:
        return env
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > (64 << 30):
            return "/dev/shm"
    except OSError:
        pass
    return "/tmp"


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
# run directories and the reference binary.  Nothing in this section imports kmdiff_b200 or maps
# libkmdiff_b200.so into this process: the synthetic rows are written by a `kmdiff-b200 dev-synth`
# subprocess that has exited before anything is timed.
# ------------------------------------------------------------------------------------------------------

def ref_binaries():
    import oracle      # constants only (paths of the reference binaries under oracle/_ref)
    return oracle.REF_BIN, oracle.REF_BIN16, oracle.REF_SMARTPCA


def make_run_dir(cfg_name: str, sample_rows: int, partitions: int) -> str:
    """kmtricks run dir holding rows [0, sample_rows) of the workload's synthetic matrix (raw matrices, or raw
    per-sample count files for the *m configs)"""
    nc, nk, k, b, rows, _ = CONFIGS[cfg_name]
    if not os.path.exists(CLI):
        raise RuntimeError(f"{CLI} not built (python -m kmdiff_b200.build)")
    d = tempfile.mkdtemp(prefix="kmdiff_run_", dir=scratch_dir())
    cmd = [CLI, "dev-synth", d, str(nc), str(nk), str(k), str(b), str(rows), str(SEED), str(ENRICHED_PPM),
           str(sample_rows), str(partitions), "counts" if cfg_name.endswith("m") else "matrix"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        shutil.rmtree(d, ignore_errors=True)
        raise RuntimeError(f"dev-synth failed ({r.returncode}): {r.stderr[-300:]}")
    return d


REF_HUNG = [0]      # reference invocations killed by the timeout (the reference has known races)


def run_reference_once(run_dir: str, cfg_name: str, threads: int, pop_args=(), keep_out: bool = False, timeout: float = 240.0):
    """one `kmdiff diff` of the unmodified reference binary -> (seconds, output dir or None).  A run that hangs is
    killed after `timeout` seconds and repeated once."""
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    ref4, ref2, _ = ref_binaries()
    binary = ref4 if b == 4 else ref2
    cmd_tail = ["-t", str(threads), "-c", corr, "-s", str(SIGNIFICANCE), "-u", str(CUTOFF)] + list(pop_args)
    last_err = ""
    for _attempt in range(2):
        out = tempfile.mkdtemp(prefix="kmdiff_ref_out_", dir=scratch_dir())
        opt = os.path.join(run_dir, "kmdiff-count.opt")
        if os.path.exists(opt) and not cfg_name.endswith("m"):
            os.remove(opt)   # reference quirk: matrix mode truncates it and the next run then fails
        cmd = [binary, "diff", "--km-run", run_dir, "-1", str(nc), "-2", str(nk), "-o", out] + cmd_tail
        t0 = time.perf_counter()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            REF_HUNG[0] += 1
            shutil.rmtree(out, ignore_errors=True)
            last_err = f"timed out after {timeout:.0f} s"
            continue
        dt = time.perf_counter() - t0
        if r.returncode != 0 or "[error]" in r.stderr:
            shutil.rmtree(out, ignore_errors=True)
            raise RuntimeError(f"reference kmdiff failed: {r.stderr[-500:]}")
        if keep_out:
            return dt, out
        shutil.rmtree(out, ignore_errors=True)
        return dt, None
    raise RuntimeError(f"reference kmdiff hung twice ({last_err})")


def reference_sample_rows(cfg_name: str, cores: int) -> int:
    nc, nk, k, b, rows, _ = CONFIGS[cfg_name]
    # ~3 M rows/s/core at N=20, scaling down with the row width; aim at a few seconds per run
    per_core = int(3_000_000 * 20 / (nc + nk))
    return int(min(rows, 32_000_000, max(200_000, per_core * min(cores, 64) * 2)))


def bench_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config or HEADLINE
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    ref4, ref2, _ = ref_binaries()
    binary = ref4 if b == 4 else ref2
    if not os.path.exists(binary):
        print(json.dumps({"impl": "reference", "unavailable": f"{binary} not built (make -C oracle refbin)"}))
        return
    cores = os.cpu_count() or 1
    sample = reference_sample_rows(cfg_name, cores)
    parts = max(cores, 8)
    run_dir = make_run_dir(cfg_name, sample, parts)
    try:
        for _ in range(args.warmup):
            run_reference_once(run_dir, cfg_name, cores)
        t = [run_reference_once(run_dir, cfg_name, cores)[0] for _ in range(args.steps)]
    finally:
        shutil.rmtree(run_dir, ignore_errors=True)
    total = sum(t)
    value = sample * args.steps / total
    line = {
        "metric": "kmer_rows_tested_per_sec", "value": value, "unit": "rows/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name), "rows_per_step": sample, "partitions": parts,
                   "row_bytes": row_bytes(cfg_name), "significance": SIGNIFICANCE, "cutoff": CUTOFF, "hung": REF_HUNG[0],
                   "note": "the unmodified reference `kmdiff diff` binary (file read + LRT + correction + FASTA) on a "
                           "bounded prefix of the workload's synthetic matrix in " + scratch_dir() + ", " +
                           ("uncompressed per-sample count files merged by km::KmerMerger" if cfg_name.endswith("m")
                            else "uncompressed matrices") + "; the run directory is written by a dev-synth subprocess "
                           "before anything is timed"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "reference",
                         "sample": f"{sample} rows in {parts} partitions per step"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def fasta_kmers(path: str):
    """the sequences of a reference FASTA file (kseq++ wraps them at 60 columns) as a set of strings"""
    out, cur = set(), []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if cur:
                    out.add("".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        out.add("".join(cur))
    return out


# ------------------------------------------------------------------------------------------------------
# our arm, per-sample count-file input (km::KmerMerger row): host streams -> device merge -> scan -> correction
# ------------------------------------------------------------------------------------------------------

def bench_merge(args, cfg_name: str | None = None, emit: bool = True, cpu_baseline: bool = True) -> dict:
    import ctypes

    import numpy as np
    import torch

    import kmdiff_b200 as K
    from kmdiff_b200 import _lib
    from kmdiff_b200 import kmtricks_io as kio

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit(f"bench.py: {args.config} is a single-GPU bench line")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — kmdiff_b200 has no CPU fallback")
    standalone = cfg_name is None
    cfg_name = cfg_name or args.config
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    if args.rows and standalone:
        rows = args.rows
    n = nc + nk
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=CONFIGS[cfg_name][4], seed=SEED, enriched_ppm=ENRICHED_PPM)
    totals = spec.totals()
    kw, rec = spec.kmer_words, 8 * spec.kmer_words + b
    # the host program's reader state, prepared once: every sample's records of a k-mer range per merge call
    # The streams live in pinned host memory, like the reader buffers of `kmdiff-b200 diff`.
    L = _lib.load()
    L.kmd_bind_host_thread(0)
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
    tree = {"algo": None, "rows": None, "keys_in": 0, "keys_out": 0}
    # the same records resident in HBM (kmd_merge_scan_device): `value`; the host streams above: `e2e`
    dev_chunks = []
    for views in chunks:
        cat = np.concatenate(views) if views else np.empty(0, np.uint8)
        dev_chunks.append((torch.from_numpy(cat).cuda(), [v.size // rec for v in views]))

    def finish():
        c = ctx.counters()
        res = ctx.correct(corr_kind, SIGNIFICANCE, c.rows)
        sig = ctx.significant()
        return res, c, sig

    def step_device(timed=False):
        ctx.reset()
        base = 0
        for d, cnt in dev_chunks:
            base += ctx.merge_scan_device(d.data_ptr(), cnt, row_base=base)
            if timed:
                m = ctx.merge_ms()
                for key in ("unpack", "passes", "rows"):
                    stage_ms[key] += m[key]
                n_passes[0] = m["n_passes"]
                info = ctx.merge_info()
                tree["algo"] = info["algo"]
                tree["rows"] = info["rows"]
                tree["keys_in"] += info["keys_in"]
                tree["keys_out"] += info["keys_out"]
                ctx.sync()
                stage_ms["scan"] += ctx.scan_kernel_ms()
        return finish()

    def step_host():
        ctx.reset()
        base = 0
        ctx.merge_prefetch(chunks[0])
        for ci, streams in enumerate(chunks):
            if ci + 1 < len(chunks):
                ctx.merge_prefetch(chunks[ci + 1])      # the next range crosses PCIe while this one is merged
            base += ctx.merge_scan_host(streams, row_base=base)
        return finish()

    def timed_loop(step):
        for _ in range(max(3, args.warmup)):
            res, c, sig = step()
        assert c.rows == rows, (c.rows, rows)
        launches0 = ctx.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res, c, sig = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, (ctx.launch_count() - launches0) // args.steps, res, c, sig

    sampler = ClockSampler(0)
    sampler.start()
    dt, launches, res, c, sig = timed_loop(step_device)
    clocks = sampler.stop()
    dt_host, _, res_h, c_h, sig_h = timed_loop(step_host)
    assert (res_h[0], res_h[1], c_h.survivors) == (res[0], res[1], c.survivors), "host streams and device records disagree"
    step_device(timed=True)               # per-stage device times, outside the timed region (adds syncs)
    value = rows * args.steps / dt
    e2e_value = rows * args.steps / dt_host
    peak, peak_src = load_peaks()
    calls = max(1, len(chunks))
    if tree["algo"] == "union":
        # every level reads the keys of its input runs and writes the keys of its output runs, once
        pass_bytes = (tree["keys_in"] + tree["keys_out"]) * 8 * kw
        kernel = "uni::union_level_kernel"
        launches_of_kernel = max(1, n_passes[0] * calls)
    else:
        pass_bytes = n_passes[0] * n_records * 2 * (8 * kw + 8)     # every pass reads and writes (key, payload)
        kernel = "mrg::merge_pass_kernel"
        launches_of_kernel = max(1, n_passes[0] * calls)
    achieved = pass_bytes / (stage_ms["passes"] * 1e-3) / 1e9 if stage_ms["passes"] else None
    # the whole device stage (keys + tree + rows + scan) against what it must move at least: the records in, the rows out and in again
    stage_total_ms = sum(stage_ms.values())
    min_bytes = n_records * rec + 2 * rows * spec.row_bytes

    cpu = None
    if cpu_baseline and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            sample = min(rows, reference_sample_rows(cfg_name[:-1], cores) // 4)
            parts = max(cores, 8)
            run_dir = make_run_dir(cfg_name, sample, parts)
            try:
                run_reference_once(run_dir, cfg_name, cores)
                dtc, _ = run_reference_once(run_dir, cfg_name, cores)
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
        "config": {"workload": workload_name(cfg_name),
                   "rows_per_gpu": rows, "records_per_gpu": n_records, "record_bytes": rec, "row_bytes": spec.row_bytes,
                   "merge_calls_per_step": len(chunks), "merge_algo": tree["algo"], "merge_rows": tree["rows"], "merge_levels": n_passes[0],
                   "level_keys_in_per_step": tree["keys_in"], "level_keys_out_per_step": tree["keys_out"],
                   "l2": f"every merge call works on {n_records // calls * rec >> 20} MiB of records and "
                         f"{n_records // calls * 8 * kw >> 20} MiB of leaf keys, above the 126 MB L2",
                   "survivors_per_gpu": int(c.survivors), "significant": [int(res[0]), int(res[1])]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None, "traffic": None, "kernel": kernel,
                     "ms_per_launch": stage_ms["passes"] / launches_of_kernel,
                     "algorithmic_bytes_per_launch": pass_bytes // launches_of_kernel, "peak_source": peak_src,
                     "stage_ms_per_step": stage_ms,
                     "device_stage": {"ms_per_step": stage_total_ms, "min_bytes_per_step": min_bytes,
                                      "gbs": min_bytes / (stage_total_ms * 1e-3) / 1e9 if stage_total_ms else None,
                                      "note": "keys + tree + rows + scan against records in + rows written and read once"}},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": n_records * rec, "d2h_bytes_per_step": int(sig.nbytes + 48),
                "rows_per_step": rows, "steps": args.steps, "ms_per_step": 1e3 * dt_host / args.steps,
                "h2d_gbs": n_records * rec * args.steps / dt_host / 1e9,
                "api": "kmd_merge_prefetch (next range's H2D on the copy stream) + kmd_merge_scan_host (host record streams) + "
                       "kmd_correct + kmd_significant; value = the same ranges through kmd_merge_scan_device, records resident in HBM"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if emit:
        print(json.dumps(line))
    ctx.close()
    del dev_chunks
    for hptr in pinned:
        L.kmd_host_free(hptr)
    return line


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

class Env:
    """process-wide state of our arm: rank, device, NCCL, stdout kept clean for the one JSON line"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — kmdiff_b200 has no CPU fallback")
        from kmdiff_b200 import _lib
        self.L = _lib.load()
        # CPU affinity + memory policy of this rank's GPU, before anything is pinned
        self.L.kmd_bind_host_thread(self.local_rank)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.saved_stdout = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout: keep fd 1 on stderr until the one JSON line is due
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs) -> list[int]:
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    def emit(self, line: dict):
        if self.rank == 0:
            if self.saved_stdout is not None:
                sys.stdout.flush()
                os.dup2(self.saved_stdout, 1)
            print(json.dumps(line), flush=True)
            if self.saved_stdout is not None:
                os.dup2(2, 1)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def h2d_ceiling_gbs(env: Env, nbytes: int = 1 << 30, reps: int = 4) -> float:
    """plain cudaMemcpyAsync from pinned memory, all ranks at once: what the PCIe path gives this rank's GPU"""
    torch = env.torch
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    dst.copy_(src, non_blocking=True)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    del src, dst
    return nbytes * reps / (ms * 1e-3) / 1e9


def measure(env: Env, args, cfg_name: str, steps: int, warmup: int, full: bool) -> dict:
    """one workload on this rank's GPU -> the measurement block (rank 0 returns the aggregate).  full = headline:
    also the CPU baseline + prefix parity and the CLI timing (N = 1, rank 0)."""
    import ctypes

    import numpy as np

    import kmdiff_b200 as K
    from kmdiff_b200.distributed import DeviceExchange, correct_sharded, global_gram
    torch = env.torch
    rank, world, dev = env.rank, env.world, env.dev
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    if args.rows and full:
        rows = args.rows
    corr_kind = K.correction_from_str(corr)
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=CONFIGS[cfg_name][4] * max(world, 1), seed=SEED,
                       enriched_ppm=ENRICHED_PPM)
    R = spec.row_bytes
    totals = spec.totals()
    thr = SIGNIFICANCE / CUTOFF

    stream = torch.cuda.current_stream()
    ctx = K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=k, count_bytes=b,
                        threshold=thr, device=env.local_rank, max_survivors=max(1 << 22, rows // 200))
    ctx.set_stream(stream.cuda_stream)
    popstrat = cfg_name == "c4"
    pop_ms = {"gram_kernel": [], "gram_call": [], "pca_host": [], "glm": [], "sampled": 0, "used": 0}
    if popstrat:
        # 1e6 sampled k-mers over ALL ranks (BASELINE.json configs[3]); every rank samples its own rows
        ctx.popstrat_enable(C4_SAMPLED / (rows * world), seed=SEED, is_diploid=True, n_pc=C4_NPC, max_sampled=2 * C4_SAMPLED)

    # ---- resident workload: this rank's rows, generated on the device (seeded, counter-based) ----
    data = torch.empty(rows * R, dtype=torch.uint8, device=dev)
    spec.rows_device(data.data_ptr(), rows, row_base=rank * rows, device=env.local_rank, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    exchange = DeviceExchange(dev)

    def pop_stage(timed: bool):
        # do_pop (cmd/diff.hpp:168-224): Gram matrix of the sampled k-mers -> PCs -> null model -> GLM per survivor
        t0 = time.perf_counter()
        # world > 1: integer Gram matrices per carrier count, NCCL all-reduce(sum), fp64 combination of the sums
        xtx, used = global_gram(ctx, dev)
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
        """-> (m, (kept control, kept case) on this rank); ONE host synchronisation (exchange_result)"""
        ctx.reset()
        ctx.scan_device(data.data_ptr(), rows, row_base=rank * rows, partition=rank)
        if popstrat:
            ctx.sync()
            last_scan_ms[0] = ctx.scan_kernel_ms()     # before the Gram kernel reuses the context's event pair
            pop_stage(True)
        return exchange.correct(ctx, corr_kind, SIGNIFICANCE)

    for _ in range(max(3, warmup)):
        total_rows, kept = step()
    launches0 = ctx.launch_count()

    sampler = ClockSampler(env.local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    env.barrier()
    ev0.record(stream)
    for _ in range(steps):
        total_rows, kept = step()
        scan_ms.append(last_scan_ms[0] if popstrat else ctx.scan_kernel_ms())
    ev1.record(stream)
    env.barrier()
    clocks = sampler.stop()
    elapsed_ms = env.max_over_ranks(ev0.elapsed_time(ev1))
    launches = (ctx.launch_count() - launches0) // steps
    value = rows * world * steps / (elapsed_ms * 1e-3)
    assert total_rows == rows * world, (total_rows, rows, world)
    dec = exchange.last
    c = ctx.counters()

    # ---- decision check (outside the timed region): the same survivors, the walk on the host ----
    check = None
    if not popstrat:
        sig_dev = ctx.significant()
        total2, kept2 = correct_sharded(ctx, corr_kind, SIGNIFICANCE, dev)
        sig_host = ctx.significant()
        same = (total2 == total_rows and tuple(kept2) == tuple(kept) and len(sig_dev) == len(sig_host) == sum(kept)
                and np.array_equal(sig_dev["row"], sig_host["row"]) and np.array_equal(sig_dev["p_value"], sig_host["p_value"]))
        n_bad = env.sum_over_ranks([0 if same else 1])[0]
        if n_bad:
            raise SystemExit(f"bench.py: device-side exchange and host walk disagree on {n_bad} rank(s): "
                             f"{(total_rows, kept)} vs {(total2, kept2)}")
        check = {"host_walk_equal": True, "compared": "m, kept control/case per rank, kept rows and p-values per rank",
                 "ranks": world}
    g_ctl, g_case, g_surv, g_cand = env.sum_over_ranks([kept[0], kept[1], c.survivors, c.candidates])

    # ---- end to end through the C-ABI with host buffers ----
    e2e_rows = min(rows, args.e2e_rows, (args.e2e_gib << 30) // R)
    L = env.L
    hptr = L.kmd_host_alloc(e2e_rows * R)
    if not hptr:
        raise SystemExit("kmd_host_alloc failed")
    hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint8)), shape=(e2e_rows * R,))
    torch.cuda.synchronize()
    hbuf[:] = data[:e2e_rows * R].cpu().numpy()      # same rows as the resident workload's prefix
    d2h = 0

    def e2e_step():
        nonlocal d2h
        ctx.reset()
        ctx.scan_host(hptr, e2e_rows, row_base=rank * rows, partition=rank)
        if popstrat:
            pop_stage(False)
        _, r = exchange.correct(ctx, corr_kind, SIGNIFICANCE)      # same global decision procedure as the resident step
        sig = ctx.significant()
        d2h = sig.nbytes + 64 + 48
        return r, len(sig)

    for _ in range(2):
        e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_res, n_sig = e2e_step()
    torch.cuda.synchronize()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_rows * world * args.e2e_steps / e2e_s
    L.kmd_host_free(hptr)
    ceiling = h2d_ceiling_gbs(env)

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

    block = {
        "value": value, "ms_per_step": elapsed_ms / steps,
        "config": {
            "workload": workload_name(cfg_name),
            "rows_per_gpu": rows, "row_bytes": R, "bytes_per_gpu": rows * R, "significance": SIGNIFICANCE,
            "cutoff": CUTOFF, "partition_sharding": f"{world} ranks x {rows} rows, disjoint",
            "l2": "inputs (>= 13 GB per step) far exceed the 126 MB L2; no flush needed",
            "correction_path": "kmdiff_b200.distributed.DeviceExchange (kmd_exchange_export -> NCCL all-gather -> "
                               "kmd_exchange_correct -> kmd_exchange_result), one host sync per step",
            "total_rows": int(total_rows), "survivors": g_surv, "screen_candidates": g_cand,
            "significant": [g_ctl, g_case], "p_cut": float(dec.p_cut), "n_accepted": int(dec.n_accepted),
            "exchange_slot_cap": exchange.cap, "exchange_retries_in_warmup": exchange.retries,
        },
        "decision_check": check,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "lrt_scan_kernel", "ms_per_launch": scan_avg_ms,
                     "algorithmic_bytes_per_launch": rows * R, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "step_frac_of_nominal_8TBs": rows * R / (elapsed_ms / steps * 1e-3) / 8e12},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": e2e_rows * R,
                "d2h_bytes_per_step": int(d2h), "rows_per_step": e2e_rows, "steps": args.e2e_steps,
                "h2d_gbs_per_gpu": e2e_rows * R * args.e2e_steps / e2e_s / 1e9,
                "h2d_ceiling_gbs_per_gpu": ceiling,
                "ceiling_note": "plain cudaMemcpyAsync from a pinned 1 GiB buffer, all ranks at once, same run",
                "api": ("kmd_scan_host + kmd_popstrat_{gram,init,apply}" if popstrat else "kmd_scan_host") +
                       " + kmd_exchange_* + kmd_significant (pinned host buffer, records read back)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if popstrat:
        n = nc + nk
        avg = lambda v: sum(v[-steps:]) / max(1, len(v[-steps:]))
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
        block["popstrat"] = {
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

    # ---- CPU baseline: the reference binary on a bounded prefix of the same matrix, and parity on that prefix ----
    if full and world == 1 and rank == 0 and not args.no_cpu_baseline:
        block.update(cpu_baseline_and_parity(env, args, cfg_name, ctx, spec, exchange, popstrat))
    elif not full and world == 1 and rank == 0 and not args.no_cpu_baseline and not popstrat:
        # the secondary workload: the same CPU baseline + prefix parity, and the drop-in program's file -> FASTA time at
        # 2e8 rows (17.6 GB) and at 64 GiB
        sec = argparse.Namespace(**vars(args))
        sec.cli_gib = f"{2e8 * R / (1 << 30)},{args.cli_gib_secondary_max}"
        block.update(cpu_baseline_and_parity(env, sec, cfg_name, ctx, spec, exchange, popstrat))

    ctx.close()
    del data
    torch.cuda.empty_cache()
    return block


def cpu_baseline_and_parity(env: Env, args, cfg_name: str, ctx, spec, exchange, popstrat: bool) -> dict:
    import numpy as np

    import kmdiff_b200 as K
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    R = spec.row_bytes
    cores = os.cpu_count() or 1
    out = {}
    ref4, ref2, smartpca = ref_binaries()
    binary = ref4 if b == 4 else ref2
    try:
        if os.path.exists(binary) and (not popstrat or os.path.exists(smartpca)):
            sample = reference_sample_rows(cfg_name, cores)
            pop_args, what = (), "`kmdiff diff"
            if popstrat:
                # the reference's whole --pop-correction pipeline (Sampler, HAWK smartpca, evec2pca.perl, GLM) with
                # the workload's sampling rate, on at most ~1 GB of rows
                sample = int(min(sample, (1 << 30) // R))
                pop_args = ("--pop-correction", "--kmer-pca", repr(C4_SAMPLED / rows), "--n-pc", str(C4_NPC), "--random-seed", str(SEED))
                what = "`kmdiff diff --pop-correction` (+ oracle/_ref/smartpca)"
            parts = max(cores, 8)
            run_dir = make_run_dir(cfg_name, sample, parts)
            ref_out = None
            try:
                run_reference_once(run_dir, cfg_name, cores, pop_args)           # warm the page cache
                dt, ref_out = run_reference_once(run_dir, cfg_name, cores, pop_args, keep_out=True)
                out["cpu_baseline"] = {
                    "value": sample / dt, "unit": "rows/s", "cores": cores, "kind": "reference", "hung": REF_HUNG[0],
                    "sample": f"{what} -t {cores}` (oracle/_ref/kmdiff) on the first {sample} rows of the workload, "
                              f"{parts} uncompressed partitions in {scratch_dir()}, {dt:.2f} s"}
                if not popstrat:
                    # parity on the prefix: the reference's FASTA k-mer set == ours (same rows, same m, same corrector)
                    ref_set = fasta_kmers(os.path.join(ref_out, "control_kmers.fasta")) | fasta_kmers(os.path.join(ref_out, "case_kmers.fasta"))
                    hrows = spec.rows_host(sample)
                    ctx.reset()
                    ctx.scan_host(hrows, row_base=0, partition=0)
                    exchange.correct(ctx, K.correction_from_str(corr), SIGNIFICANCE)      # world == 1: m = the prefix's rows
                    sig = ctx.significant()
                    ours = {K.kmer_to_string(r["kmer"][:spec.kmer_words], k) for r in sig}
                    if b != 4 or k > 32:
                        # the reference truncates k > 32 k-mers to 64 bits in its spill files (SURVEY.md §8 hazard 1)
                        ours = {s[-32:] for s in ours}
                        ref_set = {s[-32:] for s in ref_set}
                    out["parity_on_prefix"] = bool(ours == ref_set)
                    out["parity_detail"] = {"rows": sample, "reference_significant": len(ref_set), "ours_significant": len(ours),
                                            "only_reference": len(ref_set - ours), "only_ours": len(ours - ref_set)}
            finally:
                shutil.rmtree(run_dir, ignore_errors=True)
                if ref_out:
                    shutil.rmtree(ref_out, ignore_errors=True)
        else:
            import oracle
            sample = min(rows, 2_000_000)
            hrows = spec.rows_host(sample)
            m = oracle.Model(nc, nk, spec.totals())
            t0 = time.perf_counter()
            m.diff(hrows, spec.kmer_words, b, SIGNIFICANCE / CUTOFF)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": sample / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                                   "sample": f"oracle/kmd_oracle.c on the first {sample} rows, in memory, {dt:.2f} s"}
    except Exception as e:  # the baseline must never take the bench line down
        out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}
    if not popstrat and not args.no_cli:
        out["e2e_cli"] = cli_points(args, cfg_name, cores)
    return out


def cli_end_to_end(cfg_name: str, gib: float, cores: int) -> dict:
    """file -> FASTA: `kmdiff-b200 diff` vs the reference binary on the same run directory (raw matrices, page
    cache / tmpfs warm for both; best of 2 runs of the reference, of 3 of ours).  Wall clock of the whole process, start-up (CUDA context
    creation: ~0.6 s) and exit included.  A second is left between runs: the driver tears the previous process's
    GPU context and pinned memory down after it has exited, and the next CUDA initialisation would wait for that."""
    nc, nk, k, b, rows, corr = CONFIGS[cfg_name]
    R = row_bytes(cfg_name)
    n_rows = int(min(rows, gib * (1 << 30) // R))
    parts = max(cores, 8) * 2
    run_dir = make_run_dir(cfg_name, n_rows, parts)
    res = {"workload": workload_name(cfg_name), "rows": n_rows, "bytes": n_rows * R, "partitions": parts,
           "layout": "uncompressed matrices", "dir": scratch_dir(), "host_threads": cores}
    try:
        common = ["diff", "--km-run", run_dir, "-1", str(nc), "-2", str(nk), "-t", str(cores), "-c", corr,
                  "-s", str(SIGNIFICANCE), "-u", str(CUTOFF)]
        seqs = {}
        for name, binary, extra in (("reference", ref_binaries()[0 if b == 4 else 1], []),
                                    ("kmdiff_b200", CLI, ["--gpus", "1", "--count-bytes", str(b)])):
            if not os.path.exists(binary):
                continue
            best = None
            # CUDA driver + context start-up varies between 0.6 s and several seconds from one process to the next on
            # these VMs: three runs of ours, two of the reference, best of each
            for i in range(3 if name == "kmdiff_b200" else 2):
                opt = os.path.join(run_dir, "kmdiff-count.opt")
                if os.path.exists(opt):
                    os.remove(opt)
                out = tempfile.mkdtemp(prefix="kmdiff_cli_out_", dir=scratch_dir())
                time.sleep(1.0)
                t0 = time.perf_counter()
                r = subprocess.run([binary] + common + extra + ["-o", out], capture_output=True, text=True, timeout=900)
                dt = time.perf_counter() - t0
                if r.returncode != 0 or "[error]" in r.stderr:
                    shutil.rmtree(out, ignore_errors=True)
                    raise RuntimeError(f"{name} failed: {r.stderr[-300:]}")
                if i == 0:
                    seqs[name] = fasta_kmers(os.path.join(out, "control_kmers.fasta")) | fasta_kmers(os.path.join(out, "case_kmers.fasta"))
                shutil.rmtree(out, ignore_errors=True)
                best = dt if best is None else min(best, dt)
            res[name + "_s"] = best
            res[name + "_rows_per_s"] = n_rows / best
            res[name + "_file_gbs"] = n_rows * R / best / 1e9
        if "reference_s" in res and "kmdiff_b200_s" in res:
            res["speedup"] = res["reference_s"] / res["kmdiff_b200_s"]
            res["same_significant_set"] = bool(seqs["reference"] == seqs["kmdiff_b200"])
            res["significant"] = len(seqs["kmdiff_b200"])
    finally:
        shutil.rmtree(run_dir, ignore_errors=True)
    return res


def cli_points(args, cfg_name: str, cores: int) -> dict:
    """e2e_cli at every size of --cli-gib (start-up amortises with size); the largest is the headline object"""
    sizes = sorted({float(x) for x in str(args.cli_gib).split(",") if x})
    points = []
    for g in sizes:
        try:
            st = os.statvfs(scratch_dir())
            if g * (1 << 30) > 0.6 * st.f_bavail * st.f_frsize:
                points.append({"gib": g, "skipped": f"{scratch_dir()} has {st.f_bavail * st.f_frsize >> 30} GiB free"})
                continue
            points.append(cli_end_to_end(cfg_name, g, cores))
        except Exception as e:
            points.append({"gib": g, "failed": str(e)})
    done = [p for p in points if "speedup" in p]
    out = dict(done[-1]) if done else (dict(points[-1]) if points else {"failed": "no size given"})
    rest = [p for p in points if p is not (done[-1] if done else points[-1])]
    if rest:
        out["other_runs"] = rest
    return out


def bench_b200(args) -> None:
    env = Env()
    cfg_name = args.config or HEADLINE
    head = measure(env, args, cfg_name, args.steps, args.warmup, full=True)
    secondary = None
    if not args.config and not args.no_secondary and SECONDARY != cfg_name:
        secondary = measure(env, args, SECONDARY, args.steps, args.warmup, full=False)
    line = {
        "metric": "kmer_rows_tested_per_sec", "value": head["value"], "unit": "rows/s", "n_gpus": env.world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
    }
    for key in ("config", "decision_check", "roofline", "e2e", "e2e_cli", "gpu_launches", "clocks", "popstrat", "cpu_baseline",
                "parity_on_prefix", "parity_detail"):
        if key in head and head[key] is not None:
            line[key] = head[key]
    if secondary is not None:
        line[SECONDARY] = {key: secondary[key] for key in ("value", "ms_per_step", "config", "decision_check", "roofline", "e2e",
                                                           "e2e_cli", "cpu_baseline", "parity_on_prefix", "parity_detail",
                                                           "gpu_launches", "clocks") if key in secondary}
        line[SECONDARY]["unit"] = "rows/s"
    if not args.config and not args.no_secondary and env.world == 1:
        # the count-file input of the same path (km::KmerMerger row of SURVEY.md §8a): device N-way merge + scan
        try:
            env.torch.cuda.empty_cache()
            m = bench_merge(args, cfg_name="c2m", emit=False, cpu_baseline=True)
            line["c2m"] = {key: m[key] for key in ("value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches",
                                                   "cpu_baseline") if key in m}
        except (Exception, SystemExit) as e:   # this block must not take the headline line down
            line["c2m"] = {"failed": str(e)}
    env.emit(line)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="", choices=[""] + list(CONFIGS),
                    help=f"workload (default: {HEADLINE} as the headline, plus {SECONDARY} under its own key)")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the config's)")
    ap.add_argument("--e2e-rows", type=int, default=1 << 25)
    ap.add_argument("--e2e-gib", type=int, default=8, help="at most this many GiB of pinned host rows in the e2e leg")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cli-gib", default="16,64", help="size(s) in GiB of the run directories of the e2e_cli leg, comma separated")
    ap.add_argument("--cli-gib-secondary-max", type=float, default=64.0, help="largest e2e_cli run directory of the secondary (c2) workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    elif args.config.endswith("m"):
        bench_merge(args)
    else:
        bench_b200(args)


if __name__ == "__main__":
    main()
