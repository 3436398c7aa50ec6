#!/usr/bin/env python
"""f1 with numbers: how fast does the host LZ4-frame decoder (kmdiff_b200/host/lz4_frame.hpp) turn kmtricks-compressed
inputs into rows, at 1..nproc threads, against the ~55 GB/s one GPU ingests over PCIe?  Inputs: the c2 cohort as LZ4
matrices (liblz4 frames via pyarrow: default preferences = linked 64 KiB blocks, what km::lz4_stream writes) and as LZ4
per-sample count files, plus the reference's config-1 fixture (real kmtricks output)."""
import glob
import json
import os
import subprocess
import sys
import tarfile
import tempfile
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import kmdiff_b200 as K  # noqa: E402
from kmdiff_b200 import build as kb  # noqa: E402
from kmdiff_b200 import kmtricks_io as kio  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
cores = os.cpu_count() or 1
base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
res = {"cores": cores, "rows": rows, "sets": {}}
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=10**9, seed=42)
for name, count_files in (("c2 matrices .count.lz4", False), ("c2 count files .kmer.lz4", True)):
    d = tempfile.mkdtemp(prefix="lz4bench_", dir=base)
    n = rows if not count_files else rows // 4
    kio.write_run_dir(d, spec, n, max(cores, 8) * 2, lz4=True, count_files=count_files)
    files = sorted(glob.glob(os.path.join(d, "matrices", "*")) if not count_files else glob.glob(os.path.join(d, "counts", "*", "*")))
    runs = []
    for t in sorted({1, 2, 4, 8, cores}):
        if t > cores:
            continue
        out = subprocess.run([kb.CLI, "dev-lz4bench", str(t)] + files, capture_output=True, text=True, timeout=900)
        runs.append(json.loads(out.stdout))
    res["sets"][name] = {"files": len(files), "runs": runs}
    shutil.rmtree(d, ignore_errors=True)
fx = os.path.join(ROOT, "tests", "golden", "config1_run.tar.xz")
if os.path.exists(fx):
    d = tempfile.mkdtemp(prefix="lz4bench_", dir=base)
    with tarfile.open(fx) as t:
        t.extractall(d, filter="data")
    files = sorted(glob.glob(os.path.join(d, "counts", "*", "*.kmer.lz4")))
    runs = []
    for t in sorted({1, 4, cores}):
        out = subprocess.run([kb.CLI, "dev-lz4bench", str(t)] + files * 20, capture_output=True, text=True, timeout=900)
        runs.append(json.loads(out.stdout))
    res["sets"]["config 1 fixture (real kmtricks .kmer.lz4, x20)"] = {"files": len(files) * 20, "runs": runs}
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(res))
