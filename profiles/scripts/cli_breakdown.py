#!/usr/bin/env python
"""Where the wall time of one `kmdiff-b200 diff` run goes (timestamps of its own debug log), next to the
reference binary on the same run directory.  c2 shape, uncompressed matrices, page cache warm."""
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

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
gpus = sys.argv[2] if len(sys.argv) > 2 else "1"
cores = os.cpu_count() or 1
parts = max(cores, 8)
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=rows, seed=42)
d = tempfile.mkdtemp(prefix="cli_bd_run_")
kio.write_run_dir(d, spec, rows, parts)
common = ["diff", "--km-run", d, "-1", "10", "-2", "10", "-t", str(cores), "-s", "0.05", "-u", "100000"]
res = {"rows": rows, "partitions": parts, "host_threads": cores, "gpus": gpus}
for name, cmd in (("kmdiff_b200", [kb.CLI] + common + ["--gpus", gpus, "-v", "debug"]), ("reference", [oracle.REF_BIN] + common)):
    if not os.path.exists(cmd[0]):
        continue
    best, log = None, ""
    for _ in range(3):
        opt = os.path.join(d, "kmdiff-count.opt")
        if os.path.exists(opt):
            os.remove(opt)
        out = tempfile.mkdtemp(prefix="cli_bd_out_")
        t0 = time.perf_counter()
        r = subprocess.run(cmd + ["-o", out], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        shutil.rmtree(out, ignore_errors=True)
        assert "[error]" not in r.stderr, r.stderr[-400:]
        if best is None or dt < best:
            best, log = dt, r.stderr
    res[name + "_s"] = best
    res[name + "_rows_per_s"] = rows / best
    if name == "kmdiff_b200":
        res["log"] = [l for l in log.split("\n") if l][:40]
shutil.rmtree(d, ignore_errors=True)
print(json.dumps(res))
