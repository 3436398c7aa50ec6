#!/usr/bin/env python
"""Fixtures for the per-sample count-file path (km::KmerMerger), made in the build container only:

    python tests/golden/make_merge_golden.py

  km_out_dir.json     the DATA fixture of the reference's own tests/merge_test.cpp (tests/data_test/km_out_dir:
                      2 samples x 4 partitions of .kmer.lz4 files as kmtricks wrote them, the two .hist files,
                      options.txt, kmtricks.fof) as {relative path: hex of the bytes}; tests/conftest.py
                      materialises it into a temporary run directory.  merge_test.cpp expects 320 merged rows,
                      per-sample totals 160/160 and no significant k-mer.
  merge_vectors.json  dense rows produced by the UNMODIFIED km::KmerMerger (oracle/_ref,
                      ref_kmer_merge in oracle/ref_shim.cpp) on seeded per-sample streams of every
                      record shape; streams and rows stored as hex.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from kmdiff_b200 import kmtricks_io as kio  # noqa: E402

REF_FIXTURE = os.path.join(oracle.REFERENCE_ROOT, "tests", "data_test", "km_out_dir")


def seeded_streams(rng, n_samples, kmer_size, count_bytes, n_kmers, density, cmax=None):
    """per-sample record streams over a shared sorted k-mer universe"""
    kw = (kmer_size + 31) // 32
    bits = 2 * kmer_size
    universe = set()
    while len(universe) < n_kmers:
        v = int(rng.integers(0, 2**62)) | (int(rng.integers(0, 2**62)) << 62)
        universe.add(v & ((1 << bits) - 1))
    universe = sorted(universe)
    cmax = cmax or min(2**(8 * count_bytes) - 1, 50)
    streams = []
    for s in range(n_samples):
        recs = bytearray()
        for v in universe:
            if rng.random() < density:
                recs += int(v).to_bytes(8 * kw, "little")
                recs += int(rng.integers(1, cmax + 1)).to_bytes(count_bytes, "little")
        streams.append(bytes(recs))
    return streams


def main():
    files = {}
    for sub in ("counts", "histograms"):
        for r, _d, names in os.walk(os.path.join(REF_FIXTURE, sub)):
            for name in sorted(names):
                path = os.path.join(r, name)
                with open(path, "rb") as f:
                    files[os.path.relpath(path, REF_FIXTURE)] = f.read().hex()
    for name in ("options.txt", "kmtricks.fof"):
        with open(os.path.join(REF_FIXTURE, name), "rb") as f:
            files[name] = f.read().hex()
    with open(os.path.join(HERE, "km_out_dir.json"), "w") as f:
        json.dump({"source": "tlemane/kmdiff tests/data_test/km_out_dir (fixture of tests/merge_test.cpp)", "files": files}, f)

    rng = np.random.default_rng(20261017)
    cases = []
    shapes = [  # (n_samples, k, count_bytes, n_kmers, density, ab_min)
        (2, 20, 4, 60, 0.6, 1), (3, 31, 4, 80, 0.4, 1), (5, 31, 1, 100, 0.3, 2), (8, 40, 2, 90, 0.5, 3),
        (20, 31, 4, 150, 0.2, 1), (7, 63, 2, 120, 0.35, 1), (4, 32, 4, 50, 0.5, 5), (6, 21, 1, 70, 0.05, 1),
    ]
    with tempfile.TemporaryDirectory() as tmp:
        for ci, (n, k, b, m, dens, ab) in enumerate(shapes):
            streams = seeded_streams(rng, n, k, b, m, dens)
            if ci == 3:
                streams[2] = b""           # a sample without a single k-mer in this partition
            ab_vec = [ab if s % 2 == 0 else 1 for s in range(n)]
            paths = []
            for s, rec in enumerate(streams):
                p = os.path.join(tmp, f"c{ci}_s{s}.kmer.lz4")
                kio.write_kmer_file(p, np.frombuffer(rec, dtype=np.uint8), k, b, s, 0, lz4=(s % 2 == 1))
                paths.append(p)
            rows = oracle.ref_kmer_merge(paths, k, b, ab_vec)
            cases.append({"n_samples": n, "kmer_size": k, "count_bytes": b, "ab_min": ab_vec,
                          "streams": [r.hex() for r in streams], "rows": rows.tobytes().hex(), "n_rows": int(len(rows))})
    with open(os.path.join(HERE, "merge_vectors.json"), "w") as f:
        json.dump({"source": "km::KmerMerger<K, MAX_C>(paths, ab_min, k, 1, 0) of tlemane/kmdiff's vendored kmtricks, "
                             "via oracle/_ref (ref_kmer_merge)", "cases": cases}, f)
    print("wrote km_out_dir.json and merge_vectors.json:", [c["n_rows"] for c in cases])


if __name__ == "__main__":
    main()
