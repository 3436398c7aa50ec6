#!/usr/bin/env python
"""File -> FASTA wall time of `kmdiff-b200 diff` and of the reference `kmdiff diff` on the same run directory
(c2 shape: 10+10 samples, k=31, 4-byte counts).  Page cache warm for both; prints one JSON line per layout."""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import kmdiff_b200 as K  # noqa: E402
import oracle  # noqa: E402
from kmdiff_b200 import build as kb  # noqa: E402
from kmdiff_b200 import kmtricks_io as kio  # noqa: E402


def timed(cmd, run_dir):
    opt = os.path.join(run_dir, "kmdiff-count.opt")
    best = None
    for _ in range(3):
        if os.path.exists(opt):
            os.remove(opt)
        out = tempfile.mkdtemp(prefix="cli_e2e_out_")
        t0 = time.perf_counter()
        r = subprocess.run(cmd + ["-o", out], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        shutil.rmtree(out, ignore_errors=True)
        assert "[error]" not in r.stderr, r.stderr[-400:]
        best = dt if best is None else min(best, dt)
    return best


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
    cores = os.cpu_count() or 1
    parts = max(cores, 8)
    spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=rows, seed=42)
    for layout in ("matrix", "matrix_lz4", "count_files"):
        n = rows if layout != "count_files" else rows // 4
        d = tempfile.mkdtemp(prefix="cli_e2e_run_")
        kio.write_run_dir(d, spec, n, parts, lz4=(layout == "matrix_lz4"), count_files=(layout == "count_files"))
        common = ["diff", "--km-run", d, "-1", "10", "-2", "10", "-t", str(cores), "-s", "0.05", "-u", "100000"]
        ours = timed([kb.CLI] + common, d)
        ref = timed([oracle.REF_BIN] + common, d) if os.path.exists(oracle.REF_BIN) else None
        shutil.rmtree(d, ignore_errors=True)
        print(json.dumps({"layout": layout, "rows": n, "partitions": parts, "host_threads": cores,
                          "kmdiff_b200_s": ours, "kmdiff_b200_rows_per_s": n / ours,
                          "reference_s": ref, "reference_rows_per_s": n / ref if ref else None,
                          "note": "whole process wall time, best of 3 (CUDA context creation included), page cache warm"}))


if __name__ == "__main__":
    main()
